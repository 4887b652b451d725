"""Corpus feature driver on the B200 path.

Command line compatible with the reference tool (hftt_code/corpus/conv_wav2fe.py:13-50; flags -d_list, -d_wav, -d_feature, -config):

    python -m nylon_amt_b200.conv_wav2fe -d_list LISTS -d_wav WAVS -d_feature OUT -config config.json

Every name found in LISTS/{train,test,valid}.list becomes OUT/<name>.pkl, a pickle (protocol 4) of the CPU FloatTensor [T, 256]
returned by AMT.wav2feature -- the on-disk feature format the rest of the reference pipeline reads.  Launched under torchrun, each
rank takes a contiguous block of every list (one process per GPU, no collective: SURVEY.md 8e).
"""
import argparse
import json
import pathlib
import pickle

from . import shard
from .amt import AMT

SPLITS = ("train", "test", "valid")


def _names(list_dir, split):
    path = pathlib.Path(list_dir) / (split + ".list")
    if not path.is_file():
        return []
    return [line.strip() for line in path.read_text(encoding="utf-8").splitlines() if line.strip()]


def convert(list_dir, wav_dir, feature_dir, config, rank=0, world=1, log=print):
    """Returns the number of files this rank converted."""
    extractor = AMT(config, None, None)
    wav_dir, feature_dir = pathlib.Path(wav_dir), pathlib.Path(feature_dir)
    done = 0
    for split in SPLITS:
        names = _names(list_dir, split)
        lo, hi = shard.partition(len(names), world, rank)
        if rank == 0:
            log("[%s] %d files, rank 0 takes %d" % (split, len(names), hi - lo))
        for name in names[lo:hi]:
            feature = extractor.wav2feature(str(wav_dir / (name + ".wav")))
            with open(feature_dir / (name + ".pkl"), "wb") as fh:
                pickle.dump(feature, fh, protocol=4)
            done += 1
    return done


def main(argv=None):
    ap = argparse.ArgumentParser(description="wav -> log-mel feature pickles (B200)")
    for flag, text in (("-d_list", "corpus list directory"), ("-d_wav", "wav file directory (input)"),
                       ("-d_feature", "feature file directory (output)"), ("-config", "config file")):
        ap.add_argument(flag, help=text, required=True)
    ns = ap.parse_args(argv)
    rank, local_rank, world = shard.world()
    if world > 1:
        import torch
        torch.cuda.set_device(local_rank)
    config = json.loads(pathlib.Path(ns.config).read_text(encoding="utf-8"))
    n = convert(ns.d_list, ns.d_wav, ns.d_feature, config, rank, world)
    if rank == 0:
        print("conv_wav2fe: done (%d files on rank 0 of %d)" % (n, world))
    return n


if __name__ == "__main__":
    main()

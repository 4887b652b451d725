"""Evaluation-time inference driver on the B200 path: wav / feature -> activation maps -> note lists.

Command line compatible with the reference tool hftt_code/evaluation/m_inference.py:11-27 (same flags and defaults) and the same
output files: <d_mpe>/<name>_{1st,2nd}.{onset,offset,mpe,velocity} (pickle protocol 4 of the numpy maps AMT.transcript returns) and
<d_note>/<name>_{1st,2nd}.json (the note lists of AMT.mpe2note, indent 4).  Only mode 'combination' of Model_SPEC2MIDI is on the B200
hot path.  Under torchrun the file list is split into contiguous blocks per rank (one process per GPU, no collective).

    python -m nylon_amt_b200.m_inference -f_config config.json -f_list test.list -d_cp checkpoint -m best_model.pkl \
        -d_wav wav -d_fe feature -d_mpe result/mpe -d_note result/note -calc_feature -calc_transcript
"""
import argparse
import json
import pathlib
import pickle

from . import shard
from .amt import AMT

HEADS = ("onset", "offset", "mpe", "velocity")


def _dump(path, obj):
    with open(path, "wb") as fh:
        pickle.dump(obj, fh, protocol=4)


def _load(path):
    with open(path, "rb") as fh:
        return pickle.load(fh)


def run(ns, rank=0, world=1, log=print):
    config = json.loads(pathlib.Path(ns.f_config).read_text(encoding="utf-8"))
    names = [l.strip() for l in pathlib.Path(ns.f_list).read_text(encoding="utf-8").splitlines() if l.strip()]
    lo, hi = shard.partition(len(names), world, rank)
    if ns.mode != "combination" or ns.ablation:
        raise NotImplementedError("only -mode combination (no ablation) is on the B200 hot path")
    extractor = AMT(config, str(pathlib.Path(ns.d_cp) / ns.m), verbose_flag=False)
    d_wav, d_fe, d_mpe, d_note = (pathlib.Path(p) for p in (ns.d_wav, ns.d_fe, ns.d_mpe, ns.d_note))
    thresholds = dict(thred_onset=ns.thred_onset, thred_offset=ns.thred_offset, thred_mpe=ns.thred_mpe)
    for name in names[lo:hi]:
        log("[%s]" % name)
        if ns.calc_feature:
            feature = extractor.wav2feature(str(d_wav / (name + ".wav")))
            _dump(d_fe / (name + ".pkl"), feature)
        else:
            feature = _load(d_fe / (name + ".pkl"))
        if ns.calc_transcript:
            maps = extractor.transcript_stride(feature, ns.n_stride) if ns.n_stride > 0 else extractor.transcript(feature)
            for stage, four in (("1st", maps[:4]), ("2nd", maps[4:])):
                for head, arr in zip(HEADS, four):
                    _dump(d_mpe / ("%s_%s.%s" % (name, stage, head)), arr)
        else:
            maps = tuple(_load(d_mpe / ("%s_%s.%s" % (name, stage, head))) for stage in ("1st", "2nd") for head in HEADS)
        for stage, four in (("1st", maps[:4]), ("2nd", maps[4:])):
            notes = extractor.mpe2note(a_onset=four[0], a_offset=four[1], a_mpe=four[2], a_velocity=four[3], mode_velocity="ignore_zero",
                                       mode_offset="shorter", **thresholds)
            with open(d_note / ("%s_%s.json" % (name, stage)), "w", encoding="utf-8") as fh:
                json.dump(notes, fh, ensure_ascii=False, indent=4, sort_keys=False)
    return hi - lo


def main(argv=None):
    ap = argparse.ArgumentParser(description="hFT-Transformer inference for evaluation (B200)")
    ap.add_argument("-f_config", default="../corpus/config.json", help="config json file")
    ap.add_argument("-f_list", default="../corpus/MAESTRO-V3/list/test.list", help="file list")
    ap.add_argument("-d_cp", default="../checkpoint", help="checkpoint directory")
    ap.add_argument("-m", default="best_model.pkl", help="input model file")
    ap.add_argument("-mode", default="combination", help="mode to transcript (combination|single)")
    ap.add_argument("-d_wav", default="../corpus/MAESTRO-V3/wav", help="corpus wav directory")
    ap.add_argument("-d_fe", default="../corpus/MAESTRO-V3/feature", help="corpus feature directory")
    ap.add_argument("-d_mpe", default="result/mpe", help="output directory for .mpe")
    ap.add_argument("-d_note", default="result/note", help="output directory for .json")
    for flag in ("mpe", "onset", "offset"):
        ap.add_argument("-thred_" + flag, type=float, default=0.5, help="threshold value for %s detection" % flag)
    ap.add_argument("-calc_feature", action="store_true", help="flag to calculate feature data")
    ap.add_argument("-calc_transcript", action="store_true", help="flag to calculate transcript data")
    ap.add_argument("-n_stride", type=int, default=0, help="number of samples for offset")
    ap.add_argument("-ablation", action="store_true", help="ablation mode")
    ns = ap.parse_args(argv)
    rank, local_rank, world = shard.world()
    if world > 1:
        import torch
        torch.cuda.set_device(local_rank)
    n = run(ns, rank, world)
    if rank == 0:
        print("m_inference: done (%d files on rank 0 of %d)" % (n, world))
    return n


if __name__ == "__main__":
    main()

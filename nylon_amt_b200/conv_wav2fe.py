#! python
"""Driver of the feature path -- same command line as the reference's hftt_code/corpus/conv_wav2fe.py:13-50:

    python -m nylon_amt_b200.conv_wav2fe -d_list LISTS -d_wav WAVS -d_feature OUT -config config.json

For every name in {train,test,valid}.list it writes OUT/<name>.pkl holding the CPU FloatTensor [T, 256] that
AMT.wav2feature returns (pickle protocol 4, conv_wav2fe.py:46-48).  Under torchrun the files of every list are split into
contiguous blocks per rank (one process per GPU, no collective: SURVEY.md 8e); a single process converts everything.
"""
import argparse
import json
import os
import pickle

from . import amt, shard


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('-d_list', help='corpus list directory')
    parser.add_argument('-d_wav', help='wav file directory (input)')
    parser.add_argument('-d_feature', help='feature file directory (output)')
    parser.add_argument('-config', help='config file')
    args = parser.parse_args(argv)

    rank, local_rank, world = shard.world()
    if rank == 0:
        print('** conv_wav2fe: convert wav to feature **')
        print(' directory')
        print('  wav     (input) : ' + str(args.d_wav))
        print('  feature (output): ' + str(args.d_feature))
        print('  corpus list     : ' + str(args.d_list))
        print(' config file      : ' + str(args.config))

    with open(args.config, 'r', encoding='utf-8') as f:
        config = json.load(f)
    if world > 1:
        import torch
        torch.cuda.set_device(local_rank)

    AMT = amt.AMT(config, None, None)
    n_done = 0
    for attribute in ['train', 'test', 'valid']:
        path = args.d_list.rstrip('/') + '/' + str(attribute) + '.list'
        if not os.path.isfile(path):
            continue
        if rank == 0:
            print('-' + attribute + '-')
        with open(path, 'r', encoding='utf-8') as f:
            names = [l.rstrip('\n') for l in f.readlines() if l.strip()]
        lo, hi = shard.partition(len(names), world, rank)
        for fname in names[lo:hi]:
            print(fname)
            a_feature = AMT.wav2feature(args.d_wav.rstrip('/') + '/' + fname + '.wav')
            with open(args.d_feature.rstrip('/') + '/' + fname + '.pkl', 'wb') as f:
                pickle.dump(a_feature, f, protocol=4)
            n_done += 1
    if rank == 0:
        print('** done **')
    return n_done


if __name__ == '__main__':
    main()

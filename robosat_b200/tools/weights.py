"""`rs weights` (robosat/tools/weights.py:16-56): class weights `1 / ln(1.02 + p_c)` from the training masks' class histogram.

Same flag and printed output; the per-pixel counting (np.bincount over every training label) runs on the GPU
(`rsb_class_histogram`, integer-exact), PNG decoding in a thread pool."""

import argparse
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch
from PIL import Image

from robosat_b200.hostinfo import usable_cores
from robosat_b200 import _lib
from robosat_b200.config import load_config
from robosat_b200.tiles import tiles_from_slippy_map


def add_parser(subparser):
    parser = subparser.add_parser("weights", help="computes class weights on dataset", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--dataset", type=str, required=True, help="path to dataset configuration file")
    parser.set_defaults(func=main)


def class_counts_device(labels, num_classes, counts=None):
    """labels: uint8 tensor on a CUDA device; adds np.bincount(labels, minlength=C) to `counts` (uint64-as-int64 [C], device)."""
    assert labels.is_cuda and labels.dtype == torch.uint8 and labels.is_contiguous()
    if counts is None:
        counts = torch.zeros(num_classes, dtype=torch.int64, device=labels.device)
    _lib.check(_lib.load().rsb_class_histogram(labels.data_ptr(), labels.numel(), num_classes, counts.data_ptr(), _lib.current_stream_ptr()),
               "rsb_class_histogram")
    return counts


def weights_from_counts(counts, n):
    """weights.py:51-55: w = 1 / ln(1.02 + c / n), rounded to 6 decimals"""
    probs = np.asarray(counts, dtype=np.int64) / n
    weights = 1 / np.log(1.02 + probs)
    weights.round(6, out=weights)
    return weights.tolist()


def _load(path):
    return np.array(Image.open(path).convert("P"), dtype=np.uint8)  # ConvertImageMode("P") + MaskToTensor, weights.py:32


def main(args, batch=64):
    dataset = load_config(args.dataset)
    path = dataset["common"]["dataset"]
    num_classes = len(dataset["common"]["classes"])
    if not torch.cuda.is_available():
        sys.exit("Error: CUDA requested but not available")
    _lib.require_device()
    device = torch.device("cuda")
    paths = [p for _, p in sorted(tiles_from_slippy_map(os.path.join(path, "training", "labels")), key=lambda tp: tuple(int(v) for v in (tp[0].z, tp[0].x, tp[0].y)))]
    n = 0
    counts = torch.zeros(num_classes, dtype=torch.int64, device=device)
    with ThreadPoolExecutor(max_workers=min(32, usable_cores())) as pool:
        for i in range(0, len(paths), batch):
            arrays = list(pool.map(_load, paths[i:i + batch]))
            flat = torch.from_numpy(np.concatenate([a.ravel() for a in arrays]))
            n += flat.numel()
            class_counts_device(flat.to(device), num_classes, counts)
    assert n > 0, "dataset with masks must not be empty"
    host = counts.cpu().numpy()
    assert int(host.sum()) == n, "mask values outside [0, %d) found" % num_classes
    print(weights_from_counts(host, n))

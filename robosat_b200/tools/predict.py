"""`rs predict`: probability masks for slippy-map tiles -- same flags, inputs and outputs as
robosat/tools/predict.py:23-113, with the batch loop replaced by the B200 path.

Per batch the reference does H2D of fp32 tensors, DataParallel forward, softmax, D2H of B x 2 x S x S fp32, then
crop / digitize / PNG on one CPU thread. Here: the dataset yields raw uint8 tiles, `TilePredictor` normalises,
runs the U-Net plan and the softmax / crop / quantise head on the GPU and returns uint8 bins; PNG encoding runs in
a thread pool. With several GPUs the tile list is sharded by rank (one process per GPU), the checkpoint is
broadcast once over NCCL and no other collective is used.
"""

import argparse
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import torch
from PIL import Image

from robosat_b200.colors import continuous_palette_for_color
from robosat_b200.config import load_config
from robosat_b200.datasets import BufferedSlippyMapDirectory
from robosat_b200.transforms import ImageToUint8Tensor


def add_parser(subparser):
    parser = subparser.add_parser("predict", help="predicts probability masks for slippy map tiles",
                                  formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--batch_size", type=int, default=1, help="images per batch")
    parser.add_argument("--checkpoint", type=str, required=True, help="model checkpoint to load")
    parser.add_argument("--overlap", type=int, default=32, help="tile pixel overlap to predict on")
    parser.add_argument("--tile_size", type=int, required=True, help="tile size for slippy map tiles")
    parser.add_argument("--workers", type=int, default=0, help="number of workers pre-processing images")
    parser.add_argument("tiles", type=str, help="directory to read slippy map image tiles from")
    parser.add_argument("probs", type=str, help="directory to save slippy map probability masks to")
    parser.add_argument("--model", type=str, required=True, help="path to model configuration file")
    parser.add_argument("--dataset", type=str, required=True, help="path to dataset configuration file")
    parser.set_defaults(func=main)


def _save_png(root, palette, x, y, z, quantized):
    out = Image.fromarray(quantized, mode="P")
    out.putpalette(palette)
    os.makedirs(os.path.join(root, str(z), str(x)), exist_ok=True)
    out.save(os.path.join(root, str(z), str(x), str(y) + ".png"), optimize=True)


def _run(rank, world, args, port):
    from torch.utils.data import DataLoader, Subset

    from robosat_b200.dist import broadcast_state_dict, shard_range, unet_state_template
    from robosat_b200.predictor import TilePredictor

    dataset = load_config(args.dataset)
    num_classes = len(dataset["common"]["classes"])
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)

    sd = None
    if rank == 0:
        # https://github.com/pytorch/pytorch/issues/7178 -- always deserialise to host memory first
        sd = torch.load(args.checkpoint, map_location="cpu")["state_dict"]
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", str(port))
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
        sd = broadcast_state_dict(sd, unet_state_template(num_classes), device)  # the single collective of this tool

    directory = BufferedSlippyMapDirectory(args.tiles, transform=ImageToUint8Tensor(), size=args.tile_size, overlap=args.overlap)
    assert len(directory) > 0, "at least one tile in dataset"
    lo, hi = shard_range(len(directory), rank, world)
    loader = DataLoader(Subset(directory, range(lo, hi)), batch_size=args.batch_size, num_workers=args.workers)

    size = args.tile_size + 2 * args.overlap
    predictor = TilePredictor(sd, num_classes, args.batch_size, size, overlap=args.overlap, device=device)
    palette = continuous_palette_for_color("pink", 256)

    try:
        from tqdm import tqdm

        batches = tqdm(loader, desc="Eval", unit="batch", ascii=True, disable=rank != 0)
    except ImportError:  # pragma: no cover
        batches = loader

    with ThreadPoolExecutor(max_workers=max(4, (os.cpu_count() or 8) // max(world, 1))) as pool:
        pending = []
        for images, tiles in batches:
            n = images.shape[0]
            staging = predictor.pinned_input()
            staging[:n].copy_(images)
            if n < args.batch_size:
                staging[n:].zero_()  # last, ragged batch: pad with black tiles and drop their outputs
            predictor.submit(staging)
            quantized = predictor.collect().numpy()
            for tile, q in zip(tiles, quantized[:n]):
                x, y, z = (int(v) for v in tile)
                pending.append(pool.submit(_save_png, args.probs, palette, x, y, z, q.copy()))
        for f in pending:
            f.result()

    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


def main(args):
    model = load_config(args.model)
    if not model["common"]["cuda"]:
        sys.exit("Error: robosat_b200 runs on CUDA devices only; set cuda = true in the model configuration")
    if not torch.cuda.is_available():
        sys.exit("Error: CUDA requested but not available")

    world = int(os.environ.get("RSB_GPUS", torch.cuda.device_count()))
    world = max(1, min(world, torch.cuda.device_count()))
    if world == 1:
        _run(0, 1, args, 0)
    else:
        import socket

        import torch.multiprocessing as mp

        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        mp.spawn(_run, args=(world, args, port), nprocs=world, join=True)

#!/usr/bin/env python3
"""ap_calibrate_all: calibrate every raw light frame of a directory in ONE process per GPU.

Replaces the main loop of the reference's ``scripts/calibrate_all.sh`` (:353-480), which starts
``ap_calibrate.py`` once per frame (:406-411) -- four master files re-read and re-analysed for every frame.
Here the masters are loaded once and stay resident in HBM; the frames stream through
``ApCalibrate.calibrate_many`` (upload of frame k+1, fused calibrate + repair of frame k and download of frame
k-1 on three CUDA streams).  Like the shell script it writes ``<prefix><raw name>`` into the output directory
and skips frames whose output exists unless ``--clean`` is given (:383-401); ``--gpus N`` spreads the frames
over N GPUs of the box (one process each, frames rank, rank+N, ...).
"""
import argparse
import fnmatch
import logging
import os
import sys

import astrophotography_b200 as ap


def command_line_opts(argv):
    parser = argparse.ArgumentParser(
        prog="ap_calibrate_all",
        description="Bias, dark and flat calibration plus bad-pixel repair of every raw frame of a directory.")
    parser.add_argument("rawdir", metavar="RAW_DIR", help="Directory holding the raw light frames.")
    parser.add_argument("master_bias", metavar="MASTER_BIAS.FITS")
    parser.add_argument("master_dark", metavar="MASTER_DARK.FITS")
    parser.add_argument("outdir", metavar="OUT_DIR", help="Directory for the calibrated frames.")
    parser.add_argument("--master_flat", default=None, metavar="MASTER_FLAT.FITS")
    parser.add_argument("--master_badpix", default=None, metavar="MASTER_BADPIX.FITS")
    parser.add_argument("--pattern", default="*.fit*", help='Unix-style pattern of the raw frames. Default: "*.fit*"')
    parser.add_argument("--prefix", default="cal-", help='Output name prefix (calibrate_all.sh writes cal-*). Default: "cal-"')
    parser.add_argument("--deltapix", default=2, type=int, help="Half-width of the bad-pixel median box. Default: 2")
    parser.add_argument("--dark_still_biased", default=False, action="store_true",
                        help="The master dark has NOT had the bias subtracted.")
    parser.add_argument("--clean", default=False, action="store_true", help="Redo frames whose output already exists.")
    parser.add_argument("--gpus", type=int, default=1, help="GPUs of this box to spread the frames over. Default: 1")
    parser.add_argument("-l", "--loglevel", default="INFO", help="Logging message level. Default: INFO")
    return parser.parse_args(argv)


def frame_list(p):
    names = sorted(f for f in os.listdir(p.rawdir) if fnmatch.fnmatch(f, p.pattern))
    todo, skipped = [], 0
    for name in names:
        out = os.path.join(p.outdir, p.prefix + name)
        if os.path.exists(out) and not p.clean:
            skipped += 1
            continue
        todo.append((os.path.join(p.rawdir, name), out))
    return todo, skipped


def _worker(rank, world, p, todo):
    import torch
    torch.cuda.set_device(rank)
    mine = todo[rank::world]
    if not mine:
        return
    cal = ap.ApCalibrate(p.master_bias, p.master_dark, p.master_flat, p.master_badpix, p.loglevel, p.dark_still_biased)
    cal.calibrate_many([r for r, _ in mine], [o for _, o in mine], p.deltapix)


def main(args=None):
    p = command_line_opts(args)
    logger = logging.getLogger(__name__)
    os.makedirs(p.outdir, exist_ok=True)
    todo, skipped = frame_list(p)
    logger.info(f"{len(todo)} frames to calibrate, {skipped} skipped (output exists).")
    if not todo:
        return 0
    if p.gpus > 1:
        import torch.multiprocessing as mp
        mp.spawn(_worker, args=(p.gpus, p, todo), nprocs=p.gpus, join=True)
    else:
        _worker(0, 1, p, todo)
    return 0


if __name__ == "__main__":
    try:
        status = main()
    except Exception:
        logging.getLogger(__name__).critical("Shutting down due to fatal error")
        raise
    else:
        raise SystemExit(status)

"""`img2latex predict` surface (reference img2latex/cli.py:253-312) over the B200 path: same positional arguments and
option names, `Predictor.from_checkpoint` + `Predictor.predict` underneath.  The reference's console decoration
(typer / rich spinners, execution-parameter logging) is not reproduced; the two result lines are.

    python -m i2l_cli predict CHECKPOINT IMAGE [--beam-size N] [--max-length N] [--temperature T] [--top-k K] [--top-p P]
                                                [--device cuda:0] [--precision fp32|bf16]
"""
from __future__ import annotations

import argparse
import sys
from typing import List, Optional

import torch

from .predictor import Predictor


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="img2latex", description="B200-native img2latex inference")
    sub = ap.add_subparsers(dest="command", required=True)
    p = sub.add_parser("predict", help="Predict LaTeX for an image.")                       # cli.py:253-270
    p.add_argument("checkpoint_path", help="Path to trained model checkpoint")
    p.add_argument("image_path", help="Path to image file")
    p.add_argument("--beam-size", type=int, default=0, help="Beam size for beam search (0 for greedy search)")
    p.add_argument("--max-length", type=int, default=141, help="Maximum length of the generated sequence")
    p.add_argument("--temperature", type=float, default=1.0, help="Temperature for sampling")
    p.add_argument("--top-k", type=int, default=0, help="Top-k sampling parameter")
    p.add_argument("--top-p", type=float, default=0.0, help="Top-p (nucleus) sampling parameter")
    p.add_argument("--device", default=None, help="CUDA device to use (the reference's cpu / mps choices do not exist here)")
    p.add_argument("--precision", default=None, choices=["fp32", "bf16"], help="kernel set (default: I2L_PRECISION or fp32)")
    return ap


def main(argv: Optional[List[str]] = None) -> int:
    args = build_parser().parse_args(argv)
    if args.device is not None and not str(args.device).startswith("cuda"):
        raise RuntimeError(f"device {args.device!r}: this implementation is CUDA (sm_100a) only; there is no CPU / MPS path")
    device = torch.device(args.device) if args.device else None
    predictor = Predictor.from_checkpoint(args.checkpoint_path, device=device, precision=args.precision)
    latex = predictor.predict(args.image_path, beam_size=args.beam_size, max_length=args.max_length,
                              temperature=args.temperature, top_k=args.top_k, top_p=args.top_p)
    print("Generated LaTeX:")                                                               # cli.py:309-311
    print(latex)
    return 0


if __name__ == "__main__":
    sys.exit(main())

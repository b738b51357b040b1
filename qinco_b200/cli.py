"""Command line of the reference's codec, on the B200 kernels.

    python -m qinco_b200.cli --encode --model M.pt --i x.npy --o codes.npy [--raw] [--v2] [--A a --B b]
    python -m qinco_b200.cli --decode --model M.pt --i codes.npy --o y.npy [--raw] [--v2]

Same arguments and file formats as reference qinco_v1/codec_qinco.py:80-158 (`--encode/--decode --model --i --o --raw
--batch_size --device --float16`).  `--v2` reads a QINCo2 checkpoint (qinco/utils.py:118-136) instead of a pickled v1
module and, with `--o out.npz`, writes the encoded-database layout of task=encode (qinco/search/search_tasks.py:122-131).
Hydra / accelerate orchestration is out of scope; multi-GPU runs use torchrun + qinco_b200.shard.
"""
from __future__ import annotations

import argparse

import numpy as np
import torch

from . import codec, io
from .model import QINCo


def build_parser():
    ap = argparse.ArgumentParser(prog="qinco_b200.cli")
    g = ap.add_argument_group("what to do")
    g.add_argument("--encode", default=False, action="store_true")
    g.add_argument("--decode", default=False, action="store_true")
    g = ap.add_argument_group("files")
    g.add_argument("--model", default="", help="checkpoint: v1 state-dict file (or pickled module with --unsafe-pickle), v2 save_model dict with --v2")
    g.add_argument("--i", required=True, help="--encode: vectors (.npy / .fvecs / .bvecs); --decode: codes (.npy, raw, or .npz database)")
    g.add_argument("--o", required=True, help="output (npy, raw, or .npz encoded database with --v2)")
    g.add_argument("--raw", default=False, action="store_true", help="codes as headerless bit strings (M * ceil(log2 K) bits per vector)")
    g.add_argument("--v2", default=False, action="store_true", help="QINCo2 checkpoint (save_model dict) instead of a v1 module")
    g = ap.add_argument_group("computation options")
    g.add_argument("--batch_size", default=4096, type=int)
    g.add_argument("--device", default="cuda:0")
    g.add_argument("--float16", default=False, action="store_true", help="accepted for compatibility (operands are always fp16)")
    g.add_argument("--A", type=int, default=None, help="v2: override the number of pre-selected candidates")
    g.add_argument("--B", type=int, default=None, help="v2: override the beam width")
    g.add_argument("--unsafe-pickle", dest="unsafe_pickle", default=False, action="store_true",
                   help="v1: also accept a pickled nn.Module checkpoint (read through a restricted unpickler)")
    g.add_argument("--ivf_centroids", default=None, help="v2 IVF models: .npy of the (normalised) IVF centroids, like cfg.ivf_centroids")
    return ap


def main(argv=None):
    args = build_parser().parse_args(argv)
    print("arguments:", vars(args))
    assert args.encode ^ args.decode, "one of encode or decode must be selected"
    print("model file:", args.model)
    if args.v2:
        cfg, sd = io.load_v2_checkpoint(args.model, dict(A=args.A, B=args.B), ivf_centroids=args.ivf_centroids)
        ivf = bool(cfg.get("ivf_K"))
        model = QINCo(cfg, sd, device=args.device)
        M, K, D = cfg["M"], cfg["K"], cfg["D"]
    else:
        ivf = False
        model = io.load_v1_model(args.model, device=args.device, allow_pickled_module=args.unsafe_pickle)
        print("  db_scale of the model:", model.db_scale)
        M, K, D = model.M, model.K, model.D
    if args.encode:
        print("input vectors:", args.i)
        x = io.read_vectors(args.i)
        print(f"encoding {x.shape[0]} vectors of dimension {x.shape[1]}")
        if args.v2 and ivf:      # [n, M + 1]: column 0 is the IVF code, like model(batch, step="encode").T
            ivf_codes, codes_u8, _ = model._h.encode_ivf_host(np.ascontiguousarray(x, np.float32), normalize=True)
            codes = np.concatenate([ivf_codes.astype(np.int64)[:, None], codes_u8.astype(np.int64)], axis=1)
        elif args.v2:
            codes_u8, _ = model._h.encode_host(np.ascontiguousarray(x, np.float32), normalize=True)
            codes = codes_u8.astype(np.int64)
        else:
            codes = codec.encode(model, x, bs=args.batch_size, is_float16=args.float16)
        if args.raw:
            assert not ivf, "--raw packs M * ceil(log2 K) bits per vector: not defined for IVF codes"
            print(f"bit-packing codes {codes.shape}: {M} x {int(np.ceil(np.log2(K)))} bits per vector")
            io.write_raw_codes(args.o, codes, K)
        elif args.v2 and args.o.endswith(".npz"):
            io.save_encoded_db(args.o, [codes], K=K, M=M, D=D)
        else:
            print(f"writing codes {codes.shape} to {args.o}")
            np.save(args.o, codes)
    else:
        print("input codes:", args.i)
        if args.raw:
            codes = io.read_raw_codes(args.i, M, K)
        elif args.i.endswith(".npz"):
            codes, _ = io.load_encoded_db(args.i)
        elif args.i.endswith(".npy"):
            codes = np.load(args.i)
        else:
            raise RuntimeError("unrecognized format")
        print(f"decoding {codes.shape[0]} codes of {codes.shape[1]} columns")
        if args.v2 and ivf:
            y = model._h.decode_ivf_host(codes[:, 0].astype(np.int32), codes[:, 1:].astype(np.uint8), denormalize=True)
        elif args.v2:
            if codes.size and (codes.min() < 0 or codes.max() >= K):
                raise IndexError(f"codes out of range [0, {K})")
            y = model._h.decode_host(codes.astype(np.uint8), denormalize=True)
        else:
            y = codec.decode(model, codes, bs=args.batch_size, is_float16=args.float16)
        print(f"writing vectors {y.shape} to {args.o}")
        np.save(args.o, y)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()

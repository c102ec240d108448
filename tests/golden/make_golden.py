#!/usr/bin/env python
"""Generates tests/golden/golden.json from the COMPILED UPSTREAM REFERENCE (oracle/_ref, built by
oracle/Makefile from the unmodified sources under /root/reference).  Run in the authoring container:

    python tests/golden/make_golden.py

Each record names a deterministic synthetic input (generator + arguments from s2tc_b200/synth.py), the
encoder settings, the rand() cursor the call started from, and the SHA-256 of the bytes the reference
produced (tight dstRowStride).  Small cases also carry the raw output in hex.  The fixtures pin the
oracle restatement (tests/test_oracle.py) and, on the GPU box, the CUDA encoder (tests/test_gpu_golden.py)
without needing the reference sources there.
"""
import hashlib
import itertools
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402

import _oracle as O  # noqa: E402
from s2tc_b200 import synth  # noqa: E402

IMAGES = [
    ("synth_rgba", dict(width=96, height=64, seed=1234)),
    ("synth_noise", dict(width=61, height=35, seed=99)),          # ragged edges, SRGB wrap
    ("synth_normal", dict(width=64, height=64, seed=7)),
    ("synth_noise", dict(width=40, height=24, seed=3, comps=3)),  # 3-component source
    ("synth_rgba", dict(width=7, height=5, seed=5)),              # mip-tail sized
]


def make_image(gen, args):
    return getattr(synth, gen)(**args)


def main():
    assert O.ref_available(), "build oracle/_ref first (make -C oracle)"
    records = []
    for gen, args in IMAGES:
        img = make_image(gen, args)
        small = img.shape[0] * img.shape[1] <= 64
        for dxt, cd, nr, rf, di in itertools.product((0, 1, 2), range(8), (-1, 0, 7), (0, 1, 2), (0, 1, 2)):
            # keep the file small: all metrics x modes on two settings of the rest, plus a diagonal of the others
            if not ((rf == 1 and di == 1) or (rf == 2 and di == 0) or (cd == O.WAVG) or (cd == O.SRGB_MIXED and nr == 0)):
                continue
            if dxt == 2 and nr <= 0 and min(img.shape[:2]) < 4 and False:
                continue
            cursor = 17 if nr > 0 else 0
            out = O.ref_compress(img, dxt, cd, nr, rf, di, cursor=cursor)
            rec = dict(gen=gen, args=args, dxt=dxt, cd=cd, nrandom=nr, refine=rf, dither=di, cursor=cursor,
                       sha256=hashlib.sha256(out.tobytes()).hexdigest(), nbytes=int(out.size))
            if small:
                rec["hex"] = out.tobytes().hex()
            records.append(rec)
    # S3TC -> S2TC transcode vectors from the reference's convert_* routines
    transcode = []
    for dxt in (0, 1, 2):
        blocks = synth.synth_s3tc_blocks(512, dxt, seed=11)
        out = O.ref_transcode(blocks, dxt)
        transcode.append(dict(dxt=dxt, nblocks=512, seed=11, sha256=hashlib.sha256(out.tobytes()).hexdigest(),
                              first_in=blocks[:4].tobytes().hex(), first_out=out[:4 * blocks.shape[1]].tobytes().hex()))
    # pre-pass vectors
    prepass = []
    img = synth.synth_noise(53, 31, seed=21)
    for ab, di in itertools.product((1, 4, 8), (0, 1, 2)):
        out = O.ref_prepass(img, ab, di)
        prepass.append(dict(width=53, height=31, seed=21, alphabits=ab, dither=di,
                            sha256=hashlib.sha256(out.tobytes()).hexdigest()))
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(dict(source="compiled upstream reference (oracle/_ref), g++ -O3 -ffp-contract=off, glibc rand()",
                       encode=records, transcode=transcode, prepass=prepass), f, indent=0)
    print(len(records), "encode records,", len(transcode), "transcode,", len(prepass), "prepass")


if __name__ == "__main__":
    main()

"""GPU tier: the CUDA encoder against the committed golden vectors (made from the compiled upstream
reference by tests/golden/make_golden.py) and, where oracle/_ref travelled with the snapshot, against the
compiled reference itself."""
import hashlib
import json
import os

import numpy as np
import pytest

import _oracle as O
from s2tc_b200 import Settings, synth

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


def test_golden_encode_vectors(encoder):
    cache = {}
    for rec in GOLDEN["encode"]:
        key = (rec["gen"], json.dumps(rec["args"], sort_keys=True))
        if key not in cache:
            cache[key] = getattr(synth, rec["gen"])(**rec["args"])
        out = encoder.compress(cache[key], Settings(rec["dxt"], rec["cd"], rec["nrandom"], rec["refine"], rec["dither"]),
                               cursor=rec["cursor"])
        assert hashlib.sha256(out.tobytes()).hexdigest() == rec["sha256"], rec


def test_golden_prepass_and_transcode(encoder):
    for rec in GOLDEN["prepass"]:
        img = synth.synth_noise(rec["width"], rec["height"], seed=rec["seed"])
        out = encoder.rgb565_image(img, rec["alphabits"], rec["dither"])
        assert hashlib.sha256(out.tobytes()).hexdigest() == rec["sha256"], rec
    for rec in GOLDEN["transcode"]:
        blocks = synth.synth_s3tc_blocks(rec["nblocks"], rec["dxt"], seed=rec["seed"])
        out = encoder.transcode(blocks, rec["dxt"])
        assert hashlib.sha256(out.tobytes()).hexdigest() == rec["sha256"], rec


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not present")
def test_against_compiled_reference_directly(encoder):
    img = synth.synth_rgba(128, 96, seed=77)
    for dxt, cd, nr, rf, di in [(0, O.WAVG, -1, 1, 1), (2, O.SRGB_MIXED, 0, 2, 1), (0, O.WAVG, 64, 2, 0), (1, O.SRGB, 3, 1, 1),
                                (2, O.NORMALMAP, -1, 0, 0), (2, O.YUV, 12, 2, 0), (1, O.RGB, 0, 0, 1), (0, O.AVG, 0, 2, 0)]:
        got = encoder.compress(img, Settings(dxt, cd, nr, rf, di), cursor=4)
        want = O.ref_compress(img, dxt, cd, nr, rf, di, cursor=4)
        assert np.array_equal(got, want), (dxt, cd, nr, rf, di)

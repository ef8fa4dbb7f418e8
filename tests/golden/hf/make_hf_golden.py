"""Generates tests/golden/hf/*.npz from Hugging Face transformers' VitsModel (an independent implementation of the
published VITS algorithm, see tests/hf_reference.py) loaded with the synthetic *high* voice:
    python tests/golden/hf/make_hf_golden.py
Each file holds ids, scales, the Gaussian draws the forward made, the frame count and the waveform transformers
produced.  Committed so that the oracle and the CUDA path are checked against code we did not write even where
transformers is missing or a different version."""
import os
import sys
import zlib

import numpy as np
import transformers

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hf_reference as hf  # noqa: E402
from oracle import vits_oracle as vo  # noqa: E402
from sonata_b200 import voicegen  # noqa: E402


def weights_crc(t):
    c = 0
    for k in sorted(t):
        c = zlib.crc32(np.ascontiguousarray(t[k]).tobytes(), c)
    return c


def main():
    a = voicegen.ARCH["high"]
    t = voicegen.make_tensors("high")
    m = hf.load_piper_tensors(hf.build_hf_model(a), t, a)
    for n, scales, seed in ((6, (0.0, 1.0, 0.0), 0), (5, (0.667, 1.0, 0.8), 11), (4, (0.4, 1.25, 0.5), 12)):
        ids = vo.synthetic_ids(n, utt=200 + n)
        wav, ew, ez = hf.hf_infer(m, ids, *scales, seed=seed)
        noise = scales[0] != 0.0 or scales[2] != 0.0
        out = dict(ids=ids, scales=np.array(scales, np.float32), wav=wav.astype(np.float32),
                   y_len=np.int32(wav.shape[0] // int(np.prod(a["up_rates"]))), weights_crc=np.uint32(weights_crc(t)),
                   transformers_version=np.array(transformers.__version__))
        if noise:
            out["eps_w"] = np.ascontiguousarray(ew.T)      # [T, 2]      (layout of tests/golden/*.npz)
            out["eps_z"] = np.ascontiguousarray(ez.T)      # [frames, inter]
        name = f"hf_high_n{n}_{'noise' if noise else 'det'}.npz"
        np.savez_compressed(os.path.join(HERE, name), **out)
        print(name, "samples", wav.shape[0], "max|wav|", float(np.abs(wav).max()))


if __name__ == "__main__":
    main()

"""Generates the committed golden fixtures from the REFERENCE's own code (run in the build
container, where /root/reference exists; the fixtures travel, the reference does not).

  hist_kat64.raw           uint32[64][64] histogram of the 64x64 known-answer render (oracle)
  hist_big.raw             uint32[32][32] synthetic histogram with counts past 2^22 (threshold path)
  hist_zero.raw            uint32[16][16] all zero (max == 0 -> scale = inf, cudabrot.cu:436)
  tonemap_<name>_g<gamma>.pgm   output of the reference's SetGrayscalePixels + SaveImage
                           (cudabrot.cu:454-468, :548-577) on those histograms, produced by
                           oracle/_ref/ref_probe, i.e. by the reference's own host code + glibc pow

Usage: python tests/golden/make_golden.py
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

GAMMAS = ["1.0", "2.2", "0.5", "-1"]


def main():
    subprocess.run(["make", "-C", O.ORACLE_DIR], check=True)
    hists = {}
    h, _, _ = O.render(64, 64, 100, 20, 1337, 0, 1 << 16)
    hists["kat64"] = h
    rng = np.random.default_rng(20261017)
    big = rng.integers(0, 5_000_004, size=(32, 32), dtype=np.uint32)
    big[3, 5] = 5_000_003
    big[0, 0] = 0
    big[1, 1] = 1
    big[2, 2] = (1 << 22) - 1
    big[2, 3] = 1 << 22
    big[2, 4] = (1 << 22) + 1
    hists["big"] = big
    hists["zero"] = np.zeros((16, 16), dtype=np.uint32)
    for name, hist in hists.items():
        raw = os.path.join(HERE, "hist_%s.raw" % name)
        hist.astype("<u4").tofile(raw)
        for g in GAMMAS:
            out = os.path.join(HERE, "tonemap_%s_g%s.pgm" % (name, g))
            r = subprocess.run([O.REF_PROBE, "tonemap", raw, str(hist.shape[1]), str(hist.shape[0]),
                                g, out], capture_output=True, text=True)
            assert r.returncode == 0, r.stdout + r.stderr
            with open(out + ".stdout", "w") as f:
                f.write(r.stdout)
            print(name, g, r.stdout.strip())


if __name__ == "__main__":
    main()

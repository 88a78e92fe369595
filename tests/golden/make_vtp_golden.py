"""Generates tests/golden/ch4/parts_vtp.npz: the particle file written by the UNMODIFIED reference Output::particles (ch4/Output.cpp:175-229,
oracle/_ref/ref_ch4_vtp) for seeded particle sets -- the input particles and the file bytes.  Cases: num_parts < np (thinning by the
running counter), num_parts > 2 np (every second... the counter restarts at -1, so at most every second particle is written), one particle."""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_ch4_vtp")
CASES = [("thin", 5000, 400, 1), ("dense", 300, 1000, 2), ("single", 1, 10, 3), ("exact", 64, 64, 4)]


def particles(n, seed):
    rng = np.random.default_rng(seed)
    p = np.empty((7, n))
    p[:3] = rng.uniform(-0.1, 0.4, (3, n))
    p[3:6] = rng.normal(0.0, 7000.0, (3, n))
    p[6] = 1e8
    p[3, 0] = 0.0            # formatting corner cases of operator<<(double): zero, a small and a large magnitude
    if n > 2:
        p[4, 1], p[5, 2] = 1.25e-7, -3.5e11
    return p


def write_input(path, p):
    with open(path, "wb") as f:
        np.array([p.shape[1]], dtype=np.int64).tofile(f)
        np.ascontiguousarray(p).tofile(f)


def reference_file(p, num_parts):
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "results"))
        write_input(os.path.join(d, "in.bin"), p)
        subprocess.run([REF, "in.bin", str(num_parts)], cwd=d, check=True)
        return open(os.path.join(d, "results", "parts_O+_00000.vtp"), "rb").read()


if __name__ == "__main__":
    out = {}
    for name, n, num_parts, seed in CASES:
        p = particles(n, seed)
        out[name + "_part"] = p
        out[name + "_num_parts"] = np.array(num_parts)
        out[name + "_vtp"] = np.frombuffer(reference_file(p, num_parts), dtype=np.uint8)
        print(name, n, num_parts, "->", out[name + "_vtp"].size, "bytes")
    np.savez_compressed(os.path.join(HERE, "ch4", "parts_vtp.npz"), **out)

#!/usr/bin/env python
"""Writes tests/golden/inputs_v1.bin: the inputs the REFERENCE-side dumper
(rust/golden_dumper/dump_golden.rs, to be run inside the reference checkout on a box that has a
Rust toolchain) turns into tests/golden/reference_v1.bin.  Same clouds as oracle_v1.npz.

    python tests/golden/make_reference_inputs.py        # from the repo root
"""
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "oracle_v1.npz"))

CASES = [  # (kind, k, cloud[, target, target normals, max_iters])
    (1, 8, G["knn_points"]),
    (1, 10, G["normals_points"]),
    (2, 0, G["icp_src"], G["icp_tgt"], G["icp_tgt_normals"], 20),
]


def main():
    with open(os.path.join(HERE, "inputs_v1.bin"), "wb") as f:
        f.write(struct.pack("<II", 0x54433031, len(CASES)))
        for c in CASES:
            cloud = np.ascontiguousarray(c[2], "<f4")
            f.write(struct.pack("<III", c[0], c[1], len(cloud)))
            f.write(cloud.tobytes())
            if c[0] == 2:
                f.write(struct.pack("<I", len(c[3])))
                f.write(np.ascontiguousarray(c[3], "<f4").tobytes())
                f.write(np.ascontiguousarray(c[4], "<f4").tobytes())
                f.write(struct.pack("<I", c[5]))


def read_reference(path):
    """Parse reference_v1.bin (layout in rust/golden_dumper/dump_golden.rs) -> list of dicts."""
    out = []
    with open(path, "rb") as f:
        magic, n_cases = struct.unpack("<II", f.read(8))
        assert magic == 0x54433032 and n_cases == len(CASES)
        for c in CASES:
            n = len(c[2])
            if c[0] == 1:
                k1 = c[1] + 1
                idx = np.frombuffer(f.read(8 * n * k1), "<u8").reshape(n, k1)
                dist = np.frombuffer(f.read(4 * n * k1), "<f4").reshape(n, k1)
                nrm = np.frombuffer(f.read(4 * n * 6), "<f4").reshape(n, 6)
                out.append({"kind": 1, "k": c[1], "idx": idx, "dist": dist, "normals": nrm})
            else:
                T = np.frombuffer(f.read(28), "<f4")
                (mse,) = struct.unpack("<f", f.read(4))
                (iters,) = struct.unpack("<Q", f.read(8))
                (conv,) = struct.unpack("<I", f.read(4))
                (npairs,) = struct.unpack("<Q", f.read(8))
                pairs = np.frombuffer(f.read(16 * npairs), "<u8").reshape(npairs, 2)
                out.append({"kind": 2, "T": T, "mse": mse, "iterations": iters,
                            "converged": bool(conv), "pairs": pairs})
    return out


if __name__ == "__main__":
    main()

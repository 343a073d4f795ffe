"""Shared helpers for the -m gpu parity tests."""
import numpy as np


def canon(idx_row, d2_row):
    """Sort one result row by (d2 bits, index)."""
    o = np.lexsort((idx_row, d2_row))
    return idx_row[o], d2_row[o]


def knn_parity(gpu_idx, gpu_dist, ref_idx, ref_d2, pts, queries):
    """Classify every query: exact / equal modulo ties / mismatch (SURVEY §8c parity metric).
    gpu_idx u32 [nq,k]; ref_idx u64 [nq,k]; ref_d2 f32 [nq,k]."""
    nq, k = gpu_idx.shape
    gi = gpu_idx.astype(np.int64)
    ri = ref_idx.astype(np.int64)
    ri[ref_idx == np.uint64(2**64 - 1)] = 0xFFFFFFFF
    exact = np.all(gi == ri, axis=1)
    modulo, mismatch = 0, []
    for r in np.nonzero(~exact)[0]:
        a = np.sort(gi[r])
        b = np.sort(ri[r])
        if np.array_equal(a, b):
            modulo += 1  # same set, different order among equal d2
            continue
        # differing members must all be tied with the k-th distance
        diff = set(a.tolist()) ^ set(b.tolist())
        kth = ref_d2[r, -1]
        ok = True
        for j in diff:
            if j == 0xFFFFFFFF:
                ok = False
                break
            dx = pts[j].astype(np.float32) - queries[r].astype(np.float32)
            d2 = np.float32(np.float32(dx[0] * dx[0]) + np.float32(dx[1] * dx[1])) + np.float32(dx[2] * dx[2])
            if np.float32(d2) != kth:
                ok = False
        if ok:
            modulo += 1
        else:
            mismatch.append(int(r))
    return int(exact.sum()), modulo, mismatch


def angle(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    return np.arctan2(np.linalg.norm(np.cross(a, b), axis=1), (a * b).sum(1))


def quat_angle(q1, q2):
    """Rotation angle of q1 * q2^-1 (quaternions [i,j,k,w])."""
    q1 = np.asarray(q1, np.float64)
    q2 = np.asarray(q2, np.float64)
    d = abs(float(np.dot(q1, q2))) / (np.linalg.norm(q1) * np.linalg.norm(q2))
    # angle = 2 acos(d); use asin of the vector part for small angles
    v = q1[3] * (-q2[:3]) + q2[3] * q1[:3] + np.cross(q1[:3], -q2[:3])
    return 2.0 * np.arctan2(np.linalg.norm(v), d * np.linalg.norm(q1) * np.linalg.norm(q2))

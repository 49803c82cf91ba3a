"""``.rrr`` configuration files (system.c:80,213-245 reader; output.c:302-317 writer)."""
import numpy as np


def _dot(a, b):
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]


def _normalize(v):
    return v / np.sqrt(_dot(v, v))[..., None]


def orientation_from_rows(row0, row1):
    """Rows 0/1 as stored in the file -> 3x3 matrices the way system.c:225-233 builds them
    (normalise, cross product, Gram-Schmidt utils.c:91-113), operation order preserved."""
    v1 = _normalize(np.asarray(row0, dtype=np.float64))
    v2 = _normalize(np.asarray(row1, dtype=np.float64))
    v3 = np.stack([v1[..., 1] * v2[..., 2] - v1[..., 2] * v2[..., 1],
                   v1[..., 2] * v2[..., 0] - v1[..., 0] * v2[..., 2],
                   v1[..., 0] * v2[..., 1] - v1[..., 1] * v2[..., 0]], axis=-1)
    v1n2 = _dot(v1, v1)
    b1 = _dot(v2, v1) / v1n2
    v2 = v2 - b1[..., None] * v1
    v3v1, v3v2, v2n2 = _dot(v3, v1), _dot(v3, v2), _dot(v2, v2)
    c1, c2 = v3v1 / v1n2, v3v2 / v2n2
    v3 = v3 - (c1[..., None] * v1 + c2[..., None] * v2)
    return np.stack([_normalize(v1), _normalize(v2), _normalize(v3)], axis=-2)


RRB_MAGIC = b"PPRRB1\0\0"


def read_rrb(path):
    """Binary side format (host/rrr.h): magic, int64 step, int64 N, box[3], N x 3 positions, N x 9 orientation
    matrices, all little endian, full FP64.  -> (step, box[3], r[N,3], orient[N,3,3])"""
    with open(path, "rb") as f:
        if f.read(8) != RRB_MAGIC:
            raise ValueError(f"{path}: not a .rrb file")
        step, n = (int(x) for x in np.frombuffer(f.read(16), dtype="<i8"))
        box = np.frombuffer(f.read(24), dtype="<f8").copy()
        r = np.frombuffer(f.read(24 * n), dtype="<f8")
        m = np.frombuffer(f.read(72 * n), dtype="<f8")
    if r.size != 3 * n or m.size != 9 * n:
        raise ValueError(f"Number of particles found in configuration is different from the value found in the header ({n})")
    return step, box, r.reshape(n, 3).copy(), m.reshape(n, 3, 3).copy()


def write_rrb(path, step, box, r, orient):
    r = np.ascontiguousarray(r, dtype="<f8").reshape(-1, 3)
    orient = np.ascontiguousarray(orient, dtype="<f8").reshape(-1, 9)
    with open(path, "wb") as f:
        f.write(RRB_MAGIC)
        f.write(np.array([step, len(r)], dtype="<i8").tobytes())
        f.write(np.ascontiguousarray(box, dtype="<f8").tobytes())
        f.write(r.tobytes())
        f.write(orient.tobytes())


def read_rrr(path):
    """-> (step, box[3], r[N,3], orient[N,3,3]); the binary side format is recognised by its magic bytes"""
    with open(path, "rb") as f:
        if f.read(8) == RRB_MAGIC:
            return read_rrb(path)
    with open(path) as f:
        head = f.readline().split()
        step, n = int(head[0]), int(head[1])
        box = np.array([float(x) for x in head[2:5]])
        data = np.array(f.read().split(), dtype=np.float64)
    if data.size != 9 * n:
        raise ValueError(f"Number of particles found in configuration ({data.size / 9:g}) is different from the "
                         f"value found in the header ({n})")
    data = data.reshape(n, 9)
    return step, box, data[:, 6:9].copy(), orientation_from_rows(data[:, 0:3], data[:, 3:6])


def write_rrr(path, step, box, r, orient):
    """output.c:302-317: '%lld %d %.12lf %.12lf %.12lf' then rows 0/1 as %lf and position as %.12lf."""
    orient = np.asarray(orient).reshape(-1, 3, 3)
    with open(path, "w") as f:
        f.write("%d %d %.12f %.12f %.12f\n" % (step, len(r), box[0], box[1], box[2]))
        for i in range(len(r)):
            f.write("%f %f %f\n" % tuple(orient[i, 0]))
            f.write("%f %f %f\n" % tuple(orient[i, 1]))
            f.write("%.12f %.12f %.12f\n" % tuple(r[i]))

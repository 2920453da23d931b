import os
import tempfile

import numpy as np

from physx_b200 import scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    with tempfile.NamedTemporaryFile(suffix=".bin") as f:
        f.write(z["scene"].tobytes())
        f.flush()
        sc = scenes.Scene.load(f.name)
    return z, sc


def golden_order(z, t):
    return z["order"][z["order_off"][t]:z["order_off"][t + 1]]


def golden_created(z, t):
    return z["created"][z["created_off"][t]:z["created_off"][t + 1]]


def golden_deleted(z, t):
    return z["deleted"][z["deleted_off"][t]:z["deleted_off"][t + 1]]


def golden_contact_counts(z, t):
    cp = z["con_pairs"][z["con_off"][t]:z["con_off"][t + 1]]
    return {(min(int(a), int(b)), max(int(a), int(b))): int(k) for a, b, k in cp if k > 0}


def golden_contacts(z, t):
    """{(lo,hi): (k,7) array of [pos3, normal3, separation]} as reported by the reference's contact callback."""
    out = {}
    base = z["con_off"][t]
    cp = z["con_pairs"][base:z["con_off"][t + 1]]
    for i, (a, b, k) in enumerate(cp):
        if k:
            s = z["pt_off"][base + i]
            out[(min(int(a), int(b)), max(int(a), int(b)))] = (int(a), z["con_pts"][s:s + k])
    return out


def contact_counts(pairs, contacts):
    return {(int(a), int(b)): int(c[0]) for (a, b), c in zip(pairs, contacts) if c[0] > 0}


def rel_err(a, b, floor=1.0):
    """max |a-b| / max(|b|, floor): relative error with an absolute floor of `floor` scene units."""
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))

#!/usr/bin/env python
"""Pack the reference's static FLAME assets into one fp32 ``.npz`` the product path and the oracle both read.

Run in the build container (where /root/reference exists); the GPU box only sees the packed file.

What is read (reference data assets, not source):
  model_training/model/static/flame.pkl                      -> v_template, shapedirs, posedirs, J_regressor,
                                                               kintree_table, weights, f   (model/utils.py:84-89,
                                                               flame.py:124-180 say how the reference consumes them)
  model_training/model/static/indices_2d.npy                 -> indices_2d (flame.py:132)
  model_training/model/static/face_keypoints/keypoints_{191,445}/*.npy   (demo_utils.py:37-47, model_training/utils.py:81-105)
  model_training/model/static/flame_static_embedding.pkl, flame_dynamic_embedding.npy (data/utils.py:120-206)

The pickle is python-2 era and references ``chumpy.ch.Ch`` + ``scipy.sparse.csc.csc_matrix``; chumpy is not installed, so
a restricted unpickler substitutes a stub for it (its array payload is attribute ``x``).
"""
import argparse
import io
import os
import pickle
import sys
import types

import numpy as np

REF = "/root/reference"
STATIC = "model_training/model/static"


class _ChStub:
    """Stand-in for chumpy.ch.Ch: keeps whatever state the pickle gives it; payload is ``x``."""

    def __setstate__(self, state):
        self.__dict__.update(state)


class _Unpickler(pickle.Unpickler):
    _ALLOWED_PREFIX = ("numpy", "scipy.sparse", "collections", "__builtin__", "builtins", "copy_reg", "copyreg")

    def find_class(self, module, name):
        if module.startswith("chumpy"):
            return _ChStub
        if module in ("scipy.sparse.csc", "scipy.sparse._csc") and name == "csc_matrix":
            import scipy.sparse
            return scipy.sparse.csc_matrix
        if module == "__builtin__":
            module = "builtins"
        if module == "copy_reg":
            module = "copyreg"
        if not module.startswith(self._ALLOWED_PREFIX):
            raise pickle.UnpicklingError(f"refusing to unpickle {module}.{name}")
        return super().find_class(module, name)


def _load_pickle(path):
    with open(path, "rb") as f:
        return _Unpickler(f, encoding="latin1").load()


def _np(x, dtype=None):
    if isinstance(x, _ChStub):
        x = x.x
    if hasattr(x, "todense"):
        x = np.asarray(x.todense())
    x = np.asarray(x)
    return x.astype(dtype) if dtype is not None else x


def _indices_from_npy(path):
    data = np.load(path, allow_pickle=True)[()]
    out = []
    for v in data.values():
        out += list(v)
    return out


def _keypoint_subset(dirpath, exclude=()):
    names = sorted(x[:-4] for x in os.listdir(dirpath) if x.endswith(".npy"))
    idx, per_file = [], {}
    for n in names:
        if n in exclude:
            continue
        cur = _indices_from_npy(os.path.join(dirpath, n + ".npy"))
        per_file[n] = len(cur)
        idx += cur
    return np.asarray(idx, dtype=np.int32), per_file


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=REF)
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "dad_3dheads_b200", "assets",
                                                  "flame_static.npz"))
    args = ap.parse_args()
    st = os.path.join(args.ref, STATIC)

    fl = _load_pickle(os.path.join(st, "flame.pkl"))
    v_template = _np(fl["v_template"], np.float32)                      # [5023,3]
    shapedirs = _np(fl["shapedirs"], np.float32)                        # [5023,3,400]
    posedirs_raw = _np(fl["posedirs"], np.float64)                      # [5023,3,36]
    # flame.py:171-173: posedirs = reshape(posedirs, [-1, 36]).T -> [36, 15069]
    posedirs = np.reshape(posedirs_raw, [-1, posedirs_raw.shape[-1]]).T.astype(np.float32)
    j_regressor = _np(fl["J_regressor"], np.float32)                    # [5,5023]
    kintree = _np(fl["kintree_table"]).astype(np.int64)
    parents = kintree[0].copy()
    parents[0] = -1                                                      # flame.py:176-178
    weights = _np(fl["weights"], np.float32)                            # [5023,5]
    faces = _np(fl["f"]).astype(np.int32)                               # [9976,3]

    indices_2d = np.load(os.path.join(st, "indices_2d.npy")).astype(np.int32)
    kp191, files191 = _keypoint_subset(os.path.join(st, "face_keypoints", "keypoints_191"))
    kp445, files445 = _keypoint_subset(os.path.join(st, "face_keypoints", "keypoints_445"), exclude=("cheeks",))
    kp565, _ = _keypoint_subset(os.path.join(st, "face_keypoints", "keypoints_445"))

    stat = _load_pickle(os.path.join(st, "flame_static_embedding.pkl"))
    dyn = np.load(os.path.join(st, "flame_dynamic_embedding.npy"), allow_pickle=True, encoding="latin1")[()]
    static_face_idx = _np(stat["lmk_face_idx"]).astype(np.int32)         # [51]
    static_bcoords = _np(stat["lmk_b_coords"], np.float32)               # [51,3]
    dyn_face_idx = _np(dyn["lmk_face_idx"]).astype(np.int32)             # [79,17]
    dyn_bcoords = _np(dyn["lmk_b_coords"], np.float32)                   # [79,17,3]

    out = dict(
        v_template=v_template, shapedirs=shapedirs, posedirs=posedirs, J_regressor=j_regressor,
        parents=parents.astype(np.int32), lbs_weights=weights, faces=faces, indices_2d=indices_2d,
        keypoints_191=kp191, keypoints_445=kp445, keypoints_565=kp565,
        static_lmk_face_idx=static_face_idx, static_lmk_b_coords=static_bcoords,
        dynamic_lmk_face_idx=dyn_face_idx, dynamic_lmk_b_coords=dyn_bcoords,
    )
    for sub in ("head", "face", "face_w_ears", "eyeballs"):
        p = os.path.join(st, "flame_indices", sub + ".npy")
        if os.path.exists(p):
            out["flame_indices_" + sub] = np.load(p).astype(np.int32)
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    np.savez(args.out, **out)
    for k, v in out.items():
        print(f"{k:28s} {str(v.dtype):8s} {v.shape}")
    print("191 files:", files191)
    print("445 files:", files445)
    print("indices_2d == keypoints_191:", bool(np.array_equal(indices_2d, kp191)))
    print("wrote", os.path.abspath(args.out), os.path.getsize(args.out) / 1e6, "MB")


if __name__ == "__main__":
    main()

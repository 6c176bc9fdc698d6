#!/usr/bin/env python
"""Checkpoint converter for machines that have h5py (the build container does not):

    python tools/convert_weights_h5py.py to-h5  <model_type> in.npz  out.h5   [--gpus N]
    python tools/convert_weights_h5py.py to-npz <model_type> in.h5   out.npz  [--gpus N]
    python tools/convert_weights_h5py.py verify <file.h5>          # re-open a file written by minihdf5 with libhdf5

`to-h5` writes through h5py (libhdf5) instead of the package's own HDF5 writer; `verify` opens a file with h5py and
prints every group / dataset / attribute it finds -- the interoperability check that cannot run where the package was
built ("checkpoints").  --gpus N > 1 selects the multi-GPU (nested template model) layout of
l3embedding/model.py:76-77,117-119.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["to-h5", "to-npz", "verify"])
    ap.add_argument("args", nargs="+")
    ap.add_argument("--gpus", type=int, default=0)
    a = ap.parse_args()
    import h5py
    if a.cmd == "verify":
        with h5py.File(a.args[0], "r") as f:
            print("root attrs:", {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in f.attrs.items()})
            f.visititems(lambda n, o: print(n, getattr(o, "shape", ""), dict(o.attrs) if len(o.attrs) else ""))
        return
    from l3embedding_b200 import model as M
    model_type, src, dst = a.args
    m = M.load_model(src, model_type, src_num_gpus=a.gpus)     # reads either container / layout
    if a.cmd == "to-h5" and not dst.endswith((".h5", ".hdf5")):
        raise SystemExit("to-h5 needs an .h5 / .hdf5 output path")
    if a.gpus > 1:
        m = M.multi_gpu_model(m, gpus=a.gpus)
    m.save_weights(dst)                                          # weights_io uses h5py when it is importable
    print("wrote", dst)


if __name__ == "__main__":
    main()

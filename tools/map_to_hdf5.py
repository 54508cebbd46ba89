#!/usr/bin/env python
"""Convert a cora_b200 sharded map directory into the HDF5 file cora's ``write_map`` produces
(``cora/scripts/makesky.py:412-450``): dataset ``map[freq, pol, pixel]`` + ``index_map/{freq,pol,pixel}`` with the
memh5 attributes.  Needs h5py (not available in the build environment):  python tools/map_to_hdf5.py outdir out.h5"""
import json
import os
import sys

import h5py
import numpy as np

src, dst = sys.argv[1], sys.argv[2]
hdr = json.load(open(os.path.join(src, "index_map.json")))
nfreq, npol, npix = len(hdr["freq"]["centre"]), len(hdr["pol"]), hdr["npix"]
freqmap = np.zeros(nfreq, dtype=[("centre", np.float64), ("width", np.float64)])
freqmap["centre"], freqmap["width"] = hdr["freq"]["centre"], hdr["freq"]["width"]
dt = h5py.special_dtype(vlen=str)
with h5py.File(dst, "w") as f:
    f.attrs["__memh5_distributed_file"] = True
    dset = f.create_dataset("map", shape=(nfreq, npol, npix), dtype=np.float64)
    for sh in hdr["shards"]:          # one shard at a time: never more than one rank's block in memory
        dset[sh["freq_start"]:sh["freq_end"]] = np.load(os.path.join(src, sh["file"]), mmap_mode="r")
    dset.attrs["axis"] = np.array(hdr["axis"]).astype(dt)
    dset.attrs["__memh5_distributed_dset"] = True
    for name, data in (("freq", freqmap), ("pol", np.array(hdr["pol"]).astype(dt)), ("pixel", np.arange(npix))):
        d = f.create_dataset("index_map/" + name, data=data)
        d.attrs["__memh5_distributed_dset"] = False

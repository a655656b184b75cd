"""ctypes binding of the chain-output writers (csrc/gwat_chain_io.cpp) and a reader of their container.

The reference's mcmc_sampler_output writes these datasets as HDF5 (src/mcmc_io_util.cpp:555-990); the paths, shapes and the
thinning rule are the same here, the container is a flat self-describing file (no HDF5 in this image).  ``to_hdf5`` converts one
when h5py is available.
"""
import ctypes as C
import struct

import numpy as np

from . import engine

_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(t)


def write_data_dump(path, chain_ids, temperatures, positions, logl_logp=None, trim_lengths=None, ac_values=None):
    """positions [n_chains][steps][dimension]; see gwat_b200_write_data_dump (include/gwat_b200_sampler.h)."""
    lib = engine.load_library()
    pos = _f64(positions)
    n, steps, dim = pos.shape
    ids, temps, ll, trim, ac = _i32(chain_ids), _f64(temperatures), _f64(logl_logp), _i32(trim_lengths), _i32(ac_values)
    rc = lib.gwat_b200_write_data_dump(str(path).encode(), n, dim, C.c_longlong(steps), _ptr(ids, _ip), _ptr(temps, _dp), _ptr(pos, _dp),
                                       _ptr(ll, _dp), _ptr(trim, _ip), 0 if ac is None else ac.shape[0], _ptr(ac, _ip))
    if rc != 0:
        raise engine.GwatB200Error(rc, "write_data_dump failed")


def write_flat_thin_output(path, positions, ac_values, trim_lengths=None):
    """positions [n_cold][steps][dimension], ac_values [n_cold][dimension]; returns the number of rows written."""
    lib = engine.load_library()
    pos, ac, trim = _f64(positions), _i32(ac_values), _i32(trim_lengths)
    n, steps, dim = pos.shape
    rows = C.c_longlong()
    rc = lib.gwat_b200_write_flat_thin_output(str(path).encode(), n, dim, C.c_longlong(steps), _ptr(pos, _dp), _ptr(trim, _ip), _ptr(ac, _ip),
                                              C.byref(rows))
    if rc != 0:
        raise engine.GwatB200Error(rc, "write_flat_thin_output failed")
    return rows.value


def read_dump(path):
    """{dataset path: ndarray} of a container written by the functions above."""
    raw = open(path, "rb").read()
    if raw[:8] != b"GWATDUMP":
        raise ValueError("not a gwat_b200 dump: " + str(path))
    version, n = struct.unpack_from("<II", raw, 8)
    if version != 1:
        raise ValueError("unknown dump version %d" % version)
    at, out = 16, {}
    for _ in range(n):
        (ln,) = struct.unpack_from("<I", raw, at)
        name = raw[at + 4:at + 4 + ln].decode()
        at += 4 + ln
        dtype, rank = struct.unpack_from("<II", raw, at)
        dims = struct.unpack_from("<%dQ" % rank, raw, at + 8)
        at += 8 + 8 * rank
        dt = np.dtype("<f8") if dtype == 0 else np.dtype("<i4")
        count = int(np.prod(dims)) if rank else 1
        out[name] = np.frombuffer(raw, dtype=dt, count=count, offset=at).reshape(dims).copy()
        at += count * dt.itemsize
    return out


def to_hdf5(dump_path, hdf5_path):
    """Rewrite a dump as the HDF5 file the reference's tools read (needs h5py; gzip-6 chunks as in the reference)."""
    import h5py
    with h5py.File(hdf5_path, "w") as f:
        for name, arr in read_dump(dump_path).items():
            f.create_dataset(name, data=arr, compression="gzip", compression_opts=6)

"""Low-level view of one batched synthesis pass (the `sb200_job_*` entry points): used by
bench.py (device-resident timing, NCCL send buffers) and by the parity tests (noise injection,
per-stage intermediates).  Ordinary callers use `VitsModel.speak_*`."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _native as N
from .core import Audio
from .piper import _check, _take_audio


class SynthesisJob:
    def __init__(self, model, batches: Sequence[Sequence[int]], eps_w: Optional[Sequence] = None,
                 eps_z: Optional[Sequence] = None, debug: bool = False):
        self._m = model
        self._lib = model._lib
        n = len(batches)
        self.batch = n
        packed = np.ascontiguousarray(np.concatenate([np.asarray(b, dtype=np.int64) for b in batches]))
        offs = np.zeros(n + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(b) for b in batches])
        keep = []

        def ptrs(arrs):
            if arrs is None:
                return None
            out = (C.POINTER(C.c_float) * n)()
            for i, a in enumerate(arrs):
                if a is None:
                    out[i] = None
                else:
                    a = np.ascontiguousarray(a, dtype=np.float32)
                    keep.append(a)
                    out[i] = a.ctypes.data_as(C.POINTER(C.c_float))
            return out

        pw, pz = ptrs(eps_w), ptrs(eps_z)
        zf = None
        if eps_z is not None:
            zf = np.array([0 if a is None else np.asarray(a).shape[0] for a in eps_z], dtype=np.uint64)
        self._h = C.c_void_p()
        err = N.sb200_error()
        _check(self._lib.sb200_job_create(
            model._h, packed.ctypes.data_as(C.POINTER(C.c_int64)), offs.ctypes.data_as(C.POINTER(C.c_size_t)), n,
            pw, pz, None if zf is None else zf.ctypes.data_as(C.POINTER(C.c_size_t)), C.byref(self._h),
            C.byref(err)), err)
        if debug:
            self._lib.sb200_job_set_debug(self._h, 1)

    def run(self, d_out_ptr: int = 0, capacity: int = 0) -> float:
        ms, err = C.c_float(), N.sb200_error()
        _check(self._lib.sb200_job_run(self._h, C.c_void_p(d_out_ptr) if d_out_ptr else None, capacity,
                                       C.byref(ms), C.byref(err)), err)
        return float(ms.value)

    def fetch(self) -> List[Audio]:
        outs = (N.sb200_audio * self.batch)()
        err = N.sb200_error()
        _check(self._lib.sb200_job_fetch(self._h, outs, C.byref(err)), err)
        return [_take_audio(outs[i]) for i in range(self.batch)]

    def fetch_i16(self) -> List[np.ndarray]:
        """Per-utterance peak-normalised 16-bit PCM, converted on the device: bit-identical to
        `Audio.samples.to_i16_vec()` (audio/ops/src/samples.rs:51-75) at half the device->host bytes."""
        outs = (C.POINTER(C.c_int16) * self.batch)()
        lens = (C.c_size_t * self.batch)()
        err = N.sb200_error()
        _check(self._lib.sb200_job_fetch_i16(self._h, outs, lens, C.byref(err)), err)
        res = []
        for i in range(self.batch):
            n = int(lens[i])
            res.append(np.ctypeslib.as_array(outs[i], shape=(n,)).copy() if n else np.zeros(0, dtype=np.int16))
            self._lib.sb200_i16_free(outs[i])
        return res

    def copy_out(self, dst_address: int, capacity_bytes: int, fmt: int = 0) -> int:
        """Device -> host copy of the whole result (utterances back to back) into caller memory, e.g. a slice of the
        host segment shared by the ranks of one frontend; fmt 0 = f32, 1 = peak-normalised i16 PCM.  Returns bytes."""
        wr, err = C.c_size_t(), N.sb200_error()
        _check(self._lib.sb200_job_copy_out(self._h, C.c_void_p(dst_address), capacity_bytes, fmt, C.byref(wr),
                                            C.byref(err)), err)
        return int(wr.value)

    def lengths(self):
        f = (C.c_int64 * self.batch)()
        s = (C.c_int64 * self.batch)()
        o = (C.c_int64 * self.batch)()
        if self._lib.sb200_job_lengths(self._h, f, s, o) != 0:
            raise RuntimeError("job has not run")
        return list(f), list(s), list(o)

    def debug_fetch(self, name: str, b: int = 0) -> np.ndarray:
        data = C.POINTER(C.c_float)()
        rows, cols = C.c_size_t(), C.c_size_t()
        err = N.sb200_error()
        _check(self._lib.sb200_job_debug_fetch(self._h, name.encode(), b, C.byref(data), C.byref(rows),
                                               C.byref(cols), C.byref(err)), err)
        out = np.ctypeslib.as_array(data, shape=(rows.value, cols.value)).copy()
        self._lib.sb200_buffer_free(data)
        return out

    def durations(self, b: int = 0) -> np.ndarray:
        cum = C.POINTER(C.c_int32)()
        n = C.c_size_t()
        err = N.sb200_error()
        _check(self._lib.sb200_job_debug_durations(self._h, b, C.byref(cum), C.byref(n), C.byref(err)), err)
        out = np.ctypeslib.as_array(cum, shape=(n.value,)).copy()
        self._lib.sb200_buffer_free(C.cast(cum, C.POINTER(C.c_float)))
        return out

    def profile(self) -> List[dict]:
        st = (N.sb200_region_stat * 64)()
        k = self._lib.sb200_job_profile(self._h, st, 64)
        return [dict(name=st[i].name.decode(), ms=st[i].ms, flops=st[i].flops, bytes=st[i].bytes,
                     launches=st[i].launches) for i in range(k)]

    def close(self):
        if self._h:
            self._lib.sb200_job_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

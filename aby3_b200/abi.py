"""ctypes binding of include/aby3cu.h (the C ABI of libaby3cu.so).

This module is plumbing for tests and bench.py; the product host side is the
C++ sh3 facade in aby3_b200/sh3.  There is no CPU fallback: importing works
without a GPU (so symbol checks can run), but every compute call needs a B200
and raises Aby3CudaError otherwise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaby3cu.so")


class Aby3CudaError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise Aby3CudaError(
            "libaby3cu.so is missing (%s); build it with `python -m aby3_b200.build` -- "
            "there is no CPU fallback" % LIB_PATH)
    return C.CDLL(LIB_PATH)


lib = _load()

_p = C.c_void_p
_u64 = C.c_uint64
_sz = C.c_size_t
_int = C.c_int
_key = C.c_char_p

# name -> (restype, argtypes); every symbol declared in include/aby3cu.h
PROTOTYPES = {
    "aby3cu_version": (_int, []),
    "aby3cu_last_error": (C.c_char_p, []),
    "aby3cu_device_count": (_int, [C.POINTER(_int)]),
    "aby3cu_ctx_create": (_int, [_int, C.POINTER(_p)]),
    "aby3cu_ctx_create_on_stream": (_int, [_int, _p, C.POINTER(_p)]),
    "aby3cu_ctx_destroy": (_int, [_p]),
    "aby3cu_ctx_device": (_int, [_p]),
    "aby3cu_ctx_stream": (_p, [_p]),
    "aby3cu_sync": (_int, [_p]),
    "aby3cu_launch_count": (_u64, [_p]),
    "aby3cu_malloc": (_int, [_p, C.POINTER(_p), _sz]),
    "aby3cu_free": (_int, [_p, _p]),
    "aby3cu_memset": (_int, [_p, _p, _int, _sz]),
    "aby3cu_host_alloc": (_int, [C.POINTER(_p), _sz]),
    "aby3cu_host_free": (_int, [_p]),
    "aby3cu_h2d": (_int, [_p, _p, _p, _sz]),
    "aby3cu_d2h": (_int, [_p, _p, _p, _sz]),
    "aby3cu_d2d": (_int, [_p, _p, _int, _p, _int, _sz]),
    "aby3cu_event_create": (_int, [_p, C.POINTER(_p)]),
    "aby3cu_capture_begin": (_int, [_p]),
    "aby3cu_capture_end": (_int, [_p, C.POINTER(_p)]),
    "aby3cu_graph_launch": (_int, [_p, _p, _u64]),
    "aby3cu_graph_destroy": (_int, [_p]),
    "aby3cu_event_create_sync": (_int, [_p, C.POINTER(_p)]),
    "aby3cu_event_destroy": (_int, [_p]),
    "aby3cu_event_record": (_int, [_p, _p]),
    "aby3cu_event_wait": (_int, [_p, _p]),
    "aby3cu_ctx_set_corun": (_int, [_p, _int]),
    "aby3cu_trace_begin": (_int, [_p]),
    "aby3cu_trace_mark": (_int, [_p, C.c_char_p]),
    "aby3cu_trace_dump": (_int, [C.c_char_p]),
    "aby3cu_event_sync": (_int, [_p]),
    "aby3cu_event_elapsed_ms": (_int, [_p, _p, C.POINTER(C.c_float)]),
    "aby3cu_host_keystream": (_int, [_key, _u64, _sz, _p]),
    "aby3cu_aes_ctr_fill": (_int, [_p, _key, _u64, _p, _sz]),
    "aby3cu_zero_share": (_int, [_p, _key, _key, _u64, _p, _p, _sz, _int]),
    "aby3cu_mul_hadamard": (_int, [_p, _p, _p, _p, _p, _key, _key, _u64, _p, _sz]),
    "aby3cu_mul_hadamard_trunc": (_int, [_p, _p, _p, _p, _p, _key, _u64, _key, _u64, _u64, _p, _p, _p, _sz]),
    "aby3cu_trunc_tuple": (_int, [_p, _key, _u64, _key, _u64, _u64, _p, _p, _p, _p, _sz]),
    "aby3cu_trunc_tuple_at": (_int, [_p, _key, _u64, _key, _u64, _p, _u64, _u64, _p, _p, _p, _p, _sz]),
    "aby3cu_trunc_finish": (_int, [_p, _p, _p, _p, _p, _sz, _u64]),
    "aby3cu_gemm_cross": (_int, [_p, _int, _p, _p, _p, _p, _u64, _u64, _u64, _p, _int]),
    "aby3cu_gemm_cross_after": (_int, [_p, _int, _p, _p, _p, _p, _u64, _u64, _u64, _p, _int, _p]),
    "aby3cu_gemm_cross_blocks": (_int, [_p, _int, _p, _p, _p, _p, _u64, _u64, _u64, _p, _int, _p, _u64, _p, C.c_uint32]),
    "aby3cu_gemm_last_algo": (_int, [_p]),
    "aby3cu_gemm_last_main_kernel_ms": (_int, [_p, C.POINTER(C.c_float)]),
    "aby3cu_ot_send": (_int, [_p, _key, _u64, _p, _p, _sz]),
    "aby3cu_ot_help": (_int, [_p, _key, _u64, _p, _p, _sz]),
    "aby3cu_ot_recv": (_int, [_p, _p, _p, _p, _p, _sz, _int]),
    "aby3cu_bitmul_msgs_p0": (_int, [_p, _p, _p, _p, _p, _key, _u64, _key, _u64, _p, _p, _p, _sz]),
    "aby3cu_bitmul_msgs_p2": (_int, [_p, _p, _p, _p, _key, _u64, _p, _p, _sz]),
    "aby3cu_bitmul_pub_msgs": (_int, [_p, C.c_int64, _p, _p, _key, _key, _u64, _p, _sz]),
    "aby3cu_bits_expand": (_int, [_p, _p, _u64, _u64, _u64, _p]),
    "aby3cu_bitinj_msgs": (_int, [_p, _p, _p, _u64, _u64, _u64, _key, _u64, _key, _u64, _p, _p, _p]),
    "aby3cu_share_op": (_int, [_p, _int, _p, _p, _p, _sz]),
    "aby3cu_share_op2": (_int, [_p, _int, _p, _p, _p, _p, _p, _p, _sz]),
    "aby3cu_combine3": (_int, [_p, _int, _p, _p, _p, _p, _sz]),
    "aby3cu_axpb": (_int, [_p, C.c_int64, _p, C.c_int64, _p, _sz]),
    "aby3cu_transpose_i64": (_int, [_p, _p, _u64, _u64, _p]),
    "aby3cu_transpose_i64_2": (_int, [_p, _p, _p, _u64, _u64, _p, _p]),
    "aby3cu_gather_rows_multi": (_int, [_p, _int, _p, _p, _p, _p, _u64]),
    "aby3cu_gather_rows_multi_at": (_int, [_p, _int, _p, _p, _p, _p, _u64, _p]),
    "aby3cu_counter_add": (_int, [_p, _p, _u64]),
    "aby3cu_sgd_linear_colocated_work_bytes": (_sz, [_u64]),
    "aby3cu_sgd_linear_colocated": (_int, [_p, _p, _p, _p, _p, _u64, _u64, _u64, _u64, _u64, _p, _p, _p, _p, _p]),
    "aby3cu_mask_last_word": (_int, [_p, _p, _u64, _u64, _u64]),
    "aby3cu_trunc_tuple_batch_at": (_int, [_p, _int, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _u64]),
    "aby3cu_share_op_batch": (_int, [_p, _int, _int, _p, _p, _p, _sz]),
    "aby3cu_transpose_i64_batch": (_int, [_p, _int, _p, _u64, _u64, _p]),
    "aby3cu_trunc_finish_batch": (_int, [_p, _int, _p, _p, _p, _p, _sz, _u64]),
    "aby3cu_gemv_cross_batch": (_int, [_p, _int, _p, _p, _p, _p, _u64, _u64, _p, _int]),
    "aby3cu_gemv_ring": (_int, [_p, _p, _p, _p, _u64, _u64, _p, _int]),
    "aby3cu_gather_rows": (_int, [_p, _p, _u64, _p, _u64, _p]),
    "aby3cu_cmpx_gather": (_int, [_p, _p, _p, _u64, _u64, _u64, _p, _p, _p, _p]),
    "aby3cu_cmpx_scatter": (_int, [_p, _p, _p, _p, _p, _u64, _u64, _u64, _p, _p]),
    "aby3cu_iota_u64": (_int, [_p, _u64, _u64, _p, _sz]),
    "aby3cu_scatter_rows": (_int, [_p, _p, _u64, _p, _u64, _p]),
    "aby3cu_bin_row_bytes": (_u64, [_u64]),
    "aby3cu_bit_transpose": (_int, [_p, _p, _u64, _u64, _u64, _p, _u64, _p]),
    "aby3cu_bit_transpose_gather": (_int, [_p, _p, _p, _u64, _u64, _u64, _p, _u64, _p]),
    "aby3cu_bin_level": (_int, [_p, _p, C.c_uint32, _p, _p, _u64, _key, _key, _u64]),
    "aby3cu_bin_and_layer": (_int, [_p, _p, C.c_uint32, _p, _p, _u64, _key, _key, _u64]),
    "aby3cu_bin_linear_plane0": (_int, [_p, _p, C.c_uint32, _p, _p, _u64]),
    "aby3cu_bin_maxmin_rowmajor": (_int, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _u64, _u64, _key, _key, _key, _key, C.c_uint32]),
    "aby3cu_bin_bitwise_rowmajor": (_int, [_p, C.c_uint32, _p, _p, _p, _p, _p, _p, _u64, C.c_uint32, _u64, _key, _key, _u64]),
    "aby3cu_bin_pack_rows": (_int, [_p, _p, _u64, _p, C.c_uint32, _u64, _p, _p]),
    "aby3cu_bin_check_gates": (_int, [_p, _p, _p, C.c_uint32, _p, _p, _p, _u64, _u64, _p, _p]),
    "aby3cu_bin_scatter_rows": (_int, [_p, _p, _u64, _p, C.c_uint32, _u64, _p]),
}

for _name, (_res, _args) in PROTOTYPES.items():
    _f = getattr(lib, _name)          # AttributeError here == missing export
    _f.restype = _res
    _f.argtypes = _args

GEMM_AUTO, GEMM_IMAD, GEMM_TCGEN05 = 0, 1, 2
OP_ADD, OP_SUB, OP_XOR = 0, 1, 2


def check(rc):
    if rc != 0:
        raise Aby3CudaError(lib.aby3cu_last_error().decode() or "aby3cu call failed (%d)" % rc)


def device_count():
    n = _int(0)
    rc = lib.aby3cu_device_count(C.byref(n))
    return n.value if rc == 0 else 0


class DevBuf:
    """A device allocation owned by a Ctx."""

    def __init__(self, ctx, nbytes):
        self.ctx, self.nbytes = ctx, int(nbytes)
        ptr = _p()
        check(lib.aby3cu_malloc(ctx.h, C.byref(ptr), self.nbytes))
        self.ptr = ptr.value or 0

    def at(self, byte_off):
        return _p(self.ptr + int(byte_off))

    @property
    def p(self):
        return _p(self.ptr)

    def free(self):
        if self.ptr:
            check(lib.aby3cu_free(self.ctx.h, _p(self.ptr)))
            self.ptr = 0


class Ctx:
    """One party's context: device + stream."""

    def __init__(self, device=0, stream=None):
        h = _p()
        if stream is None:
            check(lib.aby3cu_ctx_create(int(device), C.byref(h)))
        else:
            check(lib.aby3cu_ctx_create_on_stream(int(device), _p(stream), C.byref(h)))
        self.h = h
        self.device = int(device)
        self._bufs = []

    def close(self):
        if self.h:
            for b in self._bufs:
                b.free()
            self._bufs = []
            lib.aby3cu_ctx_destroy(self.h)
            self.h = None

    def sync(self):
        check(lib.aby3cu_sync(self.h))

    @property
    def launches(self):
        return int(lib.aby3cu_launch_count(self.h))

    def alloc(self, nbytes):
        b = DevBuf(self, nbytes)
        self._bufs.append(b)
        return b

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        b = self.alloc(max(arr.nbytes, 16))
        if arr.nbytes:
            check(lib.aby3cu_h2d(self.h, b.p, arr.ctypes.data_as(_p), arr.nbytes))
            self.sync()
        return b

    def download(self, buf, shape, dtype=np.int64, byte_off=0):
        out = np.empty(shape, dtype=dtype)
        if out.nbytes:
            check(lib.aby3cu_d2h(self.h, out.ctypes.data_as(_p), buf.at(byte_off), out.nbytes))
            self.sync()
        return out

    def event(self):
        ev = _p()
        check(lib.aby3cu_event_create(self.h, C.byref(ev)))
        return ev

    def record(self, ev):
        check(lib.aby3cu_event_record(self.h, ev))


def elapsed_ms(start, stop):
    ms = C.c_float(0)
    check(lib.aby3cu_event_sync(stop))
    check(lib.aby3cu_event_elapsed_ms(start, stop, C.byref(ms)))
    return float(ms.value)


def host_keystream(key, byte_off, nbytes):
    out = np.empty(nbytes, dtype=np.uint8)
    check(lib.aby3cu_host_keystream(bytes(key), int(byte_off), int(nbytes), out.ctypes.data_as(_p)))
    return out

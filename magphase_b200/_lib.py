"""ctypes binding of libmagphase_b200.so (the C ABI declared in include/magphase_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is usable, every compute
entry point raises.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmagphase_b200.so')

MPB_F32, MPB_F64, MPB_I16 = 0, 1, 2
WIN_HANN, WIN_BARTLETT25 = 0, 1
_VALUE_ERRORS = (-1, -2, -3, -6)

_lib = None
_lock = threading.Lock()
_ctx = {}

_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/magphase_b200.h
SIGNATURES = {
    'mpb_create': [C.c_int, C.POINTER(_vp)],
    'mpb_destroy': [_vp],
    'mpb_last_error': [],
    'mpb_version': [],
    'mpb_launch_count': [_vp],
    'mpb_host_alloc': [_vp, _i64, C.POINTER(_vp)],
    'mpb_host_free': [_vp, _vp],
    'mpb_measure_fma_peak': [_vp, C.c_int, C.POINTER(C.c_double)],
    'mpb_profile_begin': [_vp],
    'mpb_profile_end': [_vp, C.c_char_p, _i64],
    'mpb_analysis_lossless_dev': [_vp, _vp, _vp, C.c_int, _i64, _vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int,
                                  _vp, _vp, _vp, C.c_int],
    'mpb_analysis_lossless_host': [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int, _vp, _vp, _vp],
    'mpb_frames_fft_dev': [_vp, _vp, _vp, C.c_int, _i64, _vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int, _vp, C.c_int],
    'mpb_frames_fft_host': [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int, _vp],
    'mpb_plan_ola_runs': [_vp, _vp, _i32, C.c_int, _i32, _vp, _i64, C.POINTER(_i64)],
    'mpb_synthesis_lossless_dev': [_vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _i64, _vp, _vp, _i32, _vp, _i32,
                                   C.c_int, C.c_int, _vp, C.c_int, _i64],
    'mpb_synthesis_lossless_host': [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i32, C.c_int, C.c_int, _vp, _i64],
    'mpb_mel_create': [_vp, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, C.POINTER(_vp)],
    'mpb_mel_destroy': [_vp],
    'mpb_mel_get_warp_matrix': [_vp, C.c_int, _vp],
    'mpb_mel_compress_dev': [_vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _i64, _vp, _vp, _vp, C.c_int],
    'mpb_mel_compress_host': [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp],
    'mpb_analysis_compressed_host': [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, C.c_int, _vp, _vp, _vp],
    'mpb_analysis_compressed_dev': [_vp, _vp, _vp, C.c_int, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, C.c_int],
    'mpb_sp_to_mcep_host': [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp],
    'mpb_analysis_compressed_const_hostv': [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp],
    'mpb_analysis_compressed_hostv': [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i64, C.c_int, _vp, _vp, _vp],
    'mpb_analysis_compressed_hostv2': [_vp, _vp, C.c_int, _vp, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, C.c_int],
    'mpb_syn_create': [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.POINTER(_vp)],
    'mpb_syn_destroy': [_vp],
    'mpb_synthesis_compressed_dev': [_vp, _vp, _vp, _vp, _vp, C.c_int, _i64, _vp, _vp, _i64, _vp, _vp, _i32, C.c_int,
                                     _vp, C.c_int, _i64],
    'mpb_synthesis_compressed_host': [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, C.c_int, _vp, _vp, _i64],
    'mpb_synthesis_compressed_host2': [_vp, _vp, _vp, _vp, C.c_int, _i64, _vp, _vp, _i64, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _i64],
    'mpb_post_filter_dev': [_vp, _vp, _vp, C.c_int, _i64, C.c_int, _vp, _vp, _vp, _vp],
    'mpb_post_filter_host': [_vp, _vp, _i64, C.c_int, _vp, _vp, _vp, _vp],
    'mpb_cep_energy_host': [_vp, _vp, _i64, C.c_int, _vp, C.c_int, C.c_int, _vp],
    'mpb_min_phase_dev': [_vp, _vp, _vp, C.c_int, _i64, C.c_int, _vp],
    'mpb_min_phase_host': [_vp, _vp, _i64, C.c_int, _vp],
    'mpb_mt19937_uniform_dev': [_vp, _vp, _vp, _vp, _i64, C.c_double, C.c_double, _vp, C.c_int],
    'mpb_mt19937_fill_dev': [_vp, _vp, _vp, C.c_int32, _i64, C.c_double, C.c_double, _vp, C.c_int],
    'mpb_mt19937_uniform_host': [_vp, _vp, _vp, _i64, C.c_double, C.c_double, _vp],
    'mpb_mt19937_jump_poly': [_i64, _vp],
    'mpb_analysis_geometry': [_vp, _vp, _vp, C.c_int32, _vp, C.c_double, _vp, _vp, _vp, _vp, _vp],
    'mpb_syn_geometry': [_vp, _vp, _vp, C.c_int32, C.c_int, C.c_int] + [_vp] * 11,
    'mpb_sos2_dev': [_vp, _vp, _vp, C.c_int, _vp, _i32, _vp],
    'mpb_sos2_host': [_vp, _vp, _vp, _i32, _vp],
}
_RESTYPES = {'mpb_last_error': C.c_char_p, 'mpb_version': C.c_char_p, 'mpb_launch_count': _i64}


class SynFrames(C.Structure):
    """mpb_syn_frames of include/magphase_b200.h"""
    _fields_ = [('nfrm', _i64), ('pm', _vp), ('ncentre', _vp), ('nleft', _vp), ('nright', _vp), ('voi', _vp),
                ('nkind', _vp), ('win_a', _vp), ('win_b', _vp), ('row0', _vp), ('row1', _vp), ('roww', _vp),
                ('n_utt', _i32), ('utt_frm_off', _vp), ('utt_out_off', _vp), ('utt_t0', _vp)]


def lib():
    """The loaded shared library (raises if it has not been built: run __graft_entry__.build())."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError('magphase_b200: %s is missing -- build it with '
                                       '`python -c "import __graft_entry__ as g; g.build()"`. '
                                       'There is no CPU fallback.' % LIB_PATH)
                l = C.CDLL(LIB_PATH)
                for name, args in SIGNATURES.items():
                    fn = getattr(l, name)
                    fn.argtypes = args
                    fn.restype = _RESTYPES.get(name, C.c_int)
                _lib = l
    return _lib


def check(rc):
    if rc == 0:
        return
    msg = lib().mpb_last_error().decode('utf-8', 'replace')
    if rc in _VALUE_ERRORS:
        raise ValueError('magphase_b200: %s' % msg)
    raise RuntimeError('magphase_b200 (code %d): %s' % (rc, msg))


def default_device():
    for k in ('MPB_DEVICE', 'LOCAL_RANK'):
        if os.environ.get(k, '') != '':
            return int(os.environ[k])
    return 0


_tls = threading.local()


def current_slot():
    """Context slot of the calling thread (0 unless set_thread_slot was called): every (device, slot) pair owns a private
    library context -- streams, staging buffers, scratch -- so that several host threads can drive one GPU concurrently
    (the calls of ONE context are serialised).  Plans are cached per (device, slot) as well."""
    return getattr(_tls, 'slot', 0)


def set_thread_slot(slot):
    _tls.slot = int(slot)


def ctx(device=None):
    """Context handle of (device, current_slot()) (created on first use)."""
    device = default_device() if device is None else int(device)
    key = (device, current_slot())
    with _lock:
        h = _ctx.get(key)
    if h is None:
        l = lib()
        p = _vp()
        check(l.mpb_create(device, C.byref(p)))
        with _lock:
            _ctx.setdefault(key, p)
            h = _ctx[key]
    return h


def launch_count(device=None):
    return int(lib().mpb_launch_count(ctx(device)))


def measure_fma_peak(dtype, device=None):
    """Measured non-tensor FMA peak (TFLOP/s) of the device for MPB_F32 / MPB_F64."""
    v = C.c_double()
    check(lib().mpb_measure_fma_peak(ctx(device), dtype, C.byref(v)))
    return float(v.value)


def profile_begin(device=None):
    check(lib().mpb_profile_begin(ctx(device)))


def profile_end(device=None):
    """{kernel name: (launch count, total device ms)} since profile_begin."""
    buf = C.create_string_buffer(1 << 16)
    check(lib().mpb_profile_end(ctx(device), buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.rsplit(' ', 2)
        out[name] = (int(cnt), float(ms))
    return out


class _PinnedPool:
    """Recycled page-locked result buffers.  empty(shape) hands out a NumPy array backed by pinned memory; when the
    array (and every view of it) is garbage collected the buffer goes back to the pool (bounded), so steady-state
    batch loops neither page-fault fresh memory nor bounce D2H copies through the driver's staging buffer."""
    MAX_POOLED = 2 << 30          # bytes kept for reuse
    MAX_SINGLE = 1 << 30          # larger requests use ordinary pageable memory

    def __init__(self):
        self.free = {}            # bucket size -> [address]
        self.pooled = 0
        self.lock = threading.Lock()

    @staticmethod
    def _bucket(nbytes):
        """Sizes are quantised to {1, 1.25, 1.5, 1.75} x 2^k (at most 25 % slack): batches of similar size share buffers."""
        b = 1 << 16
        while b < nbytes:
            b <<= 1
        if b <= (1 << 16):
            return b
        q = b >> 3                       # eighths of the power of two: the candidates above b / 2 are 5/8 .. 8/8 of b
        for m in (5, 6, 7, 8):
            if m * q >= nbytes:
                return m * q
        return b

    def empty(self, shape, dtype=np.float64):
        import weakref
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        if nbytes == 0 or nbytes > self.MAX_SINGLE:
            return np.empty(shape, dtype=dtype)
        size = self._bucket(nbytes)
        addr = None
        with self.lock:
            # exact bucket first, else the smallest pooled buffer that is large enough and at most twice the request
            # (page-locking a fresh buffer costs ~0.3 ms per MB: far more than carrying some slack)
            cands = [sz for sz, lst in self.free.items() if lst and size <= sz <= 2 * size]
            if cands:
                size = min(cands)
                addr = self.free[size].pop()
                self.pooled -= size
        if addr is None:
            p = _vp()
            try:
                check(lib().mpb_host_alloc(ctx(), size, C.byref(p)))
            except RuntimeError:
                return np.empty(shape, dtype=dtype)
            addr = p.value
        buf = (C.c_char * size).from_address(addr)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        weakref.finalize(buf, self._release, addr, size)
        return arr

    def _release(self, addr, size):
        with self.lock:
            if self.pooled + size <= self.MAX_POOLED:
                self.free.setdefault(size, []).append(addr)
                self.pooled += size
                return
        try:
            lib().mpb_host_free(ctx(), _vp(addr))
        except Exception:
            pass


pinned = _PinnedPool()


def ptr(a):
    """Host pointer of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(_vp)


def as_c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)

"""ctypes binding of libmagphase_b200.so (the C ABI declared in include/magphase_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is usable, every compute
entry point raises.
"""
import ctypes as C
import os
import threading
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmagphase_b200.so')

MPB_F32, MPB_F64, MPB_I16 = 0, 1, 2
WIN_HANN, WIN_BARTLETT25, WIN_RECT = 0, 1, 2
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
    'mpb_analysis_lossless_host2': [_vp, _vp, C.c_int, _i64, _vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int],
    'mpb_frames_fft_dev': [_vp, _vp, _vp, C.c_int, _i64, _vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int, _vp, C.c_int],
    'mpb_frames_fft_host': [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int, _vp],
    'mpb_plan_ola_runs': [_vp, _vp, _i32, C.c_int, _i32, _vp, _i64, C.POINTER(_i64)],
    'mpb_synthesis_lossless_dev': [_vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _i64, _vp, _vp, _i32, _vp, _i32,
                                   C.c_int, C.c_int, _vp, C.c_int, _i64],
    'mpb_synthesis_lossless_host': [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i32, C.c_int, C.c_int, _vp, _i64],
    'mpb_synthesis_lossless_host2': [_vp, _vp, _vp, _vp, C.c_int, _vp, _i64, _vp, _vp, _vp, _i32, C.c_int, C.c_int, _vp, C.c_int,
                                     _i64],
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
    'mpb_synthesis_noise_stage_dev': [_vp, _vp, _vp, _i64, _vp],
    'mpb_synthesis_compressed_host': [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, C.c_int, _vp, _vp, _i64],
    'mpb_synthesis_compressed_host2': [_vp, _vp, _vp, _vp, C.c_int, _i64, _vp, _vp, _i64, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _i64],
    'mpb_synthesis_compressed_hostv2': [_vp, _vp, _vp, _vp, _vp, _i32, C.c_int, _vp, _vp, _i64, _vp, _vp, _vp, C.c_int, _vp, _vp,
                                        C.c_int, _i64],
    'mpb_post_filter_dev': [_vp, _vp, _vp, C.c_int, _i64, C.c_int, _vp, _vp, _vp, _vp],
    'mpb_lossless_feats_host': [_vp, _vp, _i64, _vp, _vp, _vp],
    'mpb_ola_dev': [_vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int32, _vp, _i64],
    'mpb_ola_host': [_vp, _vp, _vp, _i64, C.c_int, C.c_int32, _vp, _i64],
    'mpb_post_filter_host': [_vp, _vp, _i64, C.c_int, _vp, _vp, _vp, _vp],
    'mpb_cep_energy_host': [_vp, _vp, _i64, C.c_int, _vp, C.c_int, C.c_int, _vp],
    'mpb_min_phase_dev': [_vp, _vp, _vp, C.c_int, _i64, C.c_int, _vp],
    'mpb_min_phase_host': [_vp, _vp, _i64, C.c_int, _vp],
    'mpb_mt19937_uniform_dev': [_vp, _vp, _vp, _vp, _i64, C.c_double, C.c_double, _vp, C.c_int],
    'mpb_mt19937_fill_dev': [_vp, _vp, _vp, C.c_int32, _i64, C.c_double, C.c_double, _vp, C.c_int],
    'mpb_mt19937_uniform_host': [_vp, _vp, _vp, _i64, C.c_double, C.c_double, _vp],
    'mpb_mt19937_jump_poly': [_i64, _vp],
    'mpb_analysis_geometry': [_vp, _vp, _vp, C.c_int32, _vp, C.c_double, _vp, _vp, _vp, _vp, _vp],
    'mpb_const_rate_scan': [_vp, _vp, C.c_int32, C.c_double, _vp, _vp, _vp],
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


class _NoGate:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_gate = _NoGate()


def set_device_gate(n_inflight):
    """Optional admission gate around the pipelined host entry points (None: no gate): with W worker threads and a gate of
    G < W, at most G of them are inside a C call (staging, kernels, draining) at any time; the others run their NumPy
    bookkeeping and queue at the gate.  A knob for hosts with few cores per GPU; on the development box it never beat the
    ungated run (magphase_b200.batch.run_chain_stream has the numbers).  Returns the previous gate."""
    global _gate
    prev = _gate
    _gate = _NoGate() if n_inflight is None else (n_inflight if hasattr(n_inflight, '__enter__') else threading.BoundedSemaphore(int(n_inflight)))
    return prev


def device_gate():
    return _gate


_ctx_pid = None


def ctx(device=None):
    """Context handle of (device, current_slot()) (created on first use).
    CUDA contexts do not survive fork(): a child forked after the first context was created (what the reference's
    ``lu.run_multithreaded`` would do, src/libutils.py:32-63) gets a clear error here instead of a hang or a corrupted
    device state -- the ``*_batch`` functions and ``batch.run_*`` replace that fan-out."""
    global _ctx_pid
    device = default_device() if device is None else int(device)
    key = (device, current_slot())
    with _lock:
        if _ctx and _ctx_pid != os.getpid():
            raise RuntimeError('magphase_b200: this process (pid %d) was forked after the library had initialised CUDA in pid %s; '
                               'CUDA contexts do not survive fork().  Use the *_batch functions / magphase_b200.batch instead of a '
                               'forking pool, or start the workers with the "spawn" method.' % (os.getpid(), _ctx_pid))
        h = _ctx.get(key)
    if h is None:
        l = lib()
        p = _vp()
        check(l.mpb_create(device, C.byref(p)))
        with _lock:
            _ctx.setdefault(key, p)
            _ctx_pid = os.getpid()
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
    """Page-locked result buffers, sub-allocated from a few large arenas.  empty(shape) hands out a NumPy array backed by
    pinned memory; when the array (and every view of it) is garbage collected its block goes back to the arena's free list
    (coalescing with its neighbours).  Page-locking costs ~0.3 ms per MB -- 30-50 ms for one batch's waveform -- so it must
    never happen in a steady-state batch loop, whatever the order in which worker threads take and return blocks of
    different sizes: arenas are locked once (``ARENA`` bytes each, ~0.3 s, up to ``MAX_TOTAL``) and never returned to the driver.
    Requests that do not fit (a single array larger than an arena gets an arena of its own while the budget lasts) fall
    back to ordinary pageable memory."""
    ARENA = int(os.environ.get('MPB_PINNED_ARENA_MB', '1024')) << 20
    MAX_TOTAL = int(os.environ.get('MPB_PINNED_MAX_MB', '8192')) << 20
    GRAIN = 1 << 16               # block sizes and offsets are multiples of 64 KB
    BIG = 32 << 20                # blocks from this size on are cut from the END of a free range, smaller ones from its start:
                                  # waveforms and feature matrices do not interleave, so returned ranges merge again

    def __init__(self, alloc=None):
        self.arenas = []          # [base address, size, free list [[offset, size], ...] sorted by offset]
        self.total = 0
        self.lock = threading.Lock()
        self.stats = {'allocs': 0, 'alloc_s': 0.0, 'reuses': 0, 'pageable': 0}
        self._alloc = alloc or self._cuda_alloc

    @staticmethod
    def _cuda_alloc(size):
        p = _vp()
        check(lib().mpb_host_alloc(ctx(), size, C.byref(p)))
        return p.value

    def _take(self, n):
        """Best fit over all arenas: (arena index, offset) or None.  Caller holds the lock."""
        best = None
        for ai, (_, _, free) in enumerate(self.arenas):
            for bi, (off, sz) in enumerate(free):
                if sz >= n and (best is None or sz < best[2]):
                    best = (ai, bi, sz)
        if best is None:
            return None
        ai, bi, sz = best
        free = self.arenas[ai][2]
        off = free[bi][0]
        if sz == n:
            free.pop(bi)
        elif n >= self.BIG:
            free[bi][1] = sz - n
            off += sz - n
        else:
            free[bi] = [off + n, sz - n]
        return ai, off

    def _give(self, ai, off, n):
        with self.lock:
            free = self.arenas[ai][2]
            lo, hi = 0, len(free)
            while lo < hi:
                mid = (lo + hi) // 2
                if free[mid][0] < off:
                    lo = mid + 1
                else:
                    hi = mid
            free.insert(lo, [off, n])
            if lo + 1 < len(free) and free[lo][0] + free[lo][1] == free[lo + 1][0]:
                free[lo][1] += free.pop(lo + 1)[1]
            if lo > 0 and free[lo - 1][0] + free[lo - 1][1] == free[lo][0]:
                free[lo - 1][1] += free.pop(lo)[1]

    def reserve(self, nbytes):
        """(address, release callback) of a pinned block of at least nbytes, or (None, None) when the budget is spent."""
        n = -(-int(nbytes) // self.GRAIN) * self.GRAIN
        with self.lock:
            hit = self._take(n)
            if hit is not None:
                self.stats['reuses'] += 1
            else:
                size = max(min(self.ARENA, max(64 << 20, 8 * n)), n)      # small callers lock small arenas
                if self.total + size > self.MAX_TOTAL:
                    size = n
                if self.total + size > self.MAX_TOTAL:
                    self.stats['pageable'] += 1
                    return None, None
                t0 = time.perf_counter()
                try:
                    base = self._alloc(size)
                except RuntimeError:
                    self.stats['pageable'] += 1
                    return None, None
                self.stats['allocs'] += 1
                self.stats['alloc_s'] += time.perf_counter() - t0
                self.total += size
                self.arenas.append([base, size, [[0, size]]])
                hit = self._take(n)
        ai, off = hit
        return self.arenas[ai][0] + off, (ai, off, n)

    def ensure(self, nbytes):
        """Locks arenas now until the pool holds at least nbytes (capped by MAX_TOTAL): a streaming driver that knows its
        peak demand pays for page-locking before its loop instead of at a timing-dependent moment inside it."""
        while True:
            with self.lock:
                if self.total >= min(int(nbytes), self.MAX_TOTAL) or self.total + self.ARENA > self.MAX_TOTAL:
                    return
            t0 = time.perf_counter()
            try:
                base = self._alloc(self.ARENA)
            except RuntimeError:
                return
            with self.lock:
                self.stats['allocs'] += 1
                self.stats['alloc_s'] += time.perf_counter() - t0
                self.total += self.ARENA
                self.arenas.append([base, self.ARENA, [[0, self.ARENA]]])

    def empty(self, shape, dtype=np.float64):
        import weakref
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        nbytes = count * dtype.itemsize
        if nbytes == 0:
            return np.empty(shape, dtype=dtype)
        addr, blk = self.reserve(nbytes)
        if addr is None:
            return np.empty(shape, dtype=dtype)
        buf = (C.c_char * nbytes).from_address(addr)
        arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
        weakref.finalize(buf, self._give, *blk)
        return arr


pinned = _PinnedPool()


def ptr(a):
    """Host pointer of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(_vp)


def as_c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)

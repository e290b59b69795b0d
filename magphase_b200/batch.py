"""Batch drivers: the reference's two production scripts as library functions + a CLI.

    scripts/batch_feature_extraction_for_tts.py   ->  run_feature_extraction()     wav (+ .est)  ->  .mag .real .imag .lf0 .shift
    scripts/batch_waveform_generation.py          ->  run_waveform_generation()    feature files ->  wav

The reference fans utterances out over forked processes (``lu.run_multithreaded``, src/libutils.py:32-63), each doing
its own file IO and NumPy work.  Here the GPU does the arithmetic for a whole batch per call, so the host's job is to
keep it fed: a pool of IO threads reads (and decodes) the files of batch k+1 and writes the results of batch k-1 while
batch k is on the device -- disk, PCIe and kernels overlap.  File formats are the reference's own (SURVEY section 8(f)
rank 2): PCM wav, raw little-endian float32 rows without header (src/libutils.py:112-127), REAPER ``.est`` text.

Results are the same files the per-utterance wrappers write (``analysis_for_acoustic_modelling``,
``synthesis_from_acoustic_modelling``): feature files byte for byte; waveforms from the same NumPy noise stream, consumed
in list order exactly like the reference's sequential loop (``b_multiproc = False``, its default for synthesis).

Two things a 100k-utterance job needs that the reference's scripts leave to the operator (failure
detection / checkpoint-resume rows): ``on_error='skip'`` isolates a token whose files cannot be read or whose marks are
unusable -- it is appended to ``crash_file_list_<host>_<pid>.scp`` in the output directory (the convention of
scripts/batch_convert_label_state_aligned_to_variable_frame_rate.py:48, 59-70) and the batch goes on without it; and
``resume=True`` skips tokens whose output files are already complete (outputs are written under a temporary name and
renamed, so a killed job never leaves a half-written file that looks finished).  A resumed generation run still draws the
skipped utterances' share of the NumPy noise stream, so its waveforms are the ones an uninterrupted run writes.

    python -m magphase_b200.batch extract  --scp file_id.scp --wav-dir wavs --out-dir feats [--est-dir est] [--resume] [--skip-errors]
    python -m magphase_b200.batch generate --scp file_id.scp --feats-dir feats --out-dir wavs_syn --fs 48000 [--resume] [--skip-errors]
"""
import argparse
import concurrent.futures as cf
import os
import socket
import time
import warnings

import numpy as np

from . import hostio as io
from . import magphase as mp


def read_tokens(files_scp):
    """One file token per line, '#' comments (lu.read_text_file2, src/libutils.py:92-96)."""
    with open(files_scp) as f:
        return [ln.split('#')[0].strip() for ln in f if ln.split('#')[0].strip()]


def _batches(items, n):
    for a in range(0, len(items), n):
        yield items[a:a + n]


def _load_utterance(in_wav_dir, est_dir, token):
    wav = os.path.join(in_wav_dir, token + '.wav')
    v_sig, fs = io.read_audio_file(wav)
    est = os.path.join(est_dir, token + '.est') if est_dir else None
    v_pm_sec, v_voi = mp.get_pitch_marks_and_voicing(wav, len(v_sig), fs, est_file=est)
    return v_sig, fs, v_pm_sec, v_voi


def _feature_exts(b_const_rate):
    return ('.mag', '.real', '.imag', '.lf0') + (() if b_const_rate else ('.shift',))


def _atomically(write, path, *args):
    """write(..., tmp) then rename: the final name only ever holds a complete file (what ``resume`` relies on).  The
    temporary name keeps the extension (sf.write / wavfile pick the container from it)."""
    d, base = os.path.split(path)
    tmp = os.path.join(d, '.tmp%d_%s' % (os.getpid(), base))
    try:
        write(*args, tmp)
        os.replace(tmp, path)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)


def _write_features(out_dir, token, feats, b_const_rate):
    for ext, arr in zip(_feature_exts(b_const_rate), feats[:5]):
        _atomically(io.write_binfile, os.path.join(out_dir, token + ext), arr)


def _write_wav(path, y, fs):
    _atomically(lambda y_, fs_, tmp: io.write_audio_file(tmp, y_, fs_), path, y, fs)


def _complete(paths):
    return all(os.path.isfile(p) and os.path.getsize(p) > 0 for p in paths)


class _CrashList:
    """Tokens that failed, appended (and flushed) one per line to crash_file_list_<host>_<pid>.scp as they happen."""

    def __init__(self, out_dir):
        self.path = os.path.join(out_dir, 'crash_file_list_%s_%d.scp' % (socket.gethostname(), os.getpid()))
        self.tokens = []

    def add(self, token, err):
        self.tokens.append(token)
        warnings.warn('%s: %s: %s (skipped, listed in %s)' % (token, type(err).__name__, err, self.path))
        with open(self.path, 'a') as f:
            f.write('%s # %s: %s\n' % (token, type(err).__name__, str(err).replace('\n', ' ')))


def _collect(futures, tokens, on_error, crashes):
    """Results of the per-token loader futures; with on_error='skip' a failing token goes to the crash list instead."""
    ok_tokens, items = [], []
    for t, f in zip(tokens, futures):
        try:
            items.append(f.result())
            ok_tokens.append(t)
        except Exception as e:                      # noqa: BLE001 -- any per-utterance failure is isolated, as the reference's label script does
            if on_error != 'skip':
                raise
            crashes.add(t, e)
    return ok_tokens, items


def _check_on_error(on_error):
    if on_error not in ('raise', 'skip'):
        raise ValueError("on_error must be 'raise' or 'skip'")


def run_feature_extraction(tokens, in_wav_dir, out_feats_dir, est_dir=None, fft_len=None, mag_dim=60, phase_dim=45,
                           b_const_rate=False, batch_utts=64, io_threads=8, verbose=False, resume=False, on_error='raise'):
    """analysis_for_acoustic_modelling (src/magphase.py:2992-3022) for a list of file tokens (or an .scp path).
    Pitch marks come from ``<est_dir>/<token>.est`` when est_dir is given, else from the REAPER binary (as the reference).
    resume / on_error: module docstring.  Returns {'utterances', 'frames', 'seconds', 'skipped', 'failed'}."""
    _check_on_error(on_error)
    if isinstance(tokens, str):
        tokens = read_tokens(tokens)
    tokens = list(tokens)
    os.makedirs(out_feats_dir, exist_ok=True)
    skipped = []
    if resume:
        done = lambda t: _complete([os.path.join(out_feats_dir, t + e) for e in _feature_exts(b_const_rate)])
        skipped = [t for t in tokens if done(t)]
        gone = set(skipped)
        tokens = [t for t in tokens if t not in gone]
    crashes = _CrashList(out_feats_dir)
    t0 = time.perf_counter()
    n_frames = n_done = 0
    with cf.ThreadPoolExecutor(max_workers=io_threads) as pool:
        groups = list(_batches(tokens, batch_utts))
        load = lambda g: [pool.submit(_load_utterance, in_wav_dir, est_dir, t) for t in g]
        pending = load(groups[0]) if groups else []
        writes = []
        for k, g in enumerate(groups):
            g, utts = _collect(pending, g, on_error, crashes)
            pending = load(groups[k + 1]) if k + 1 < len(groups) else []      # disk reads overlap the GPU call below
            if not utts:
                continue
            fs = utts[0][1]
            if any(u[1] != fs for u in utts):
                raise ValueError('all utterances of a batch must share the sample rate')
            call = lambda uu: mp.analysis_compressed_batch(
                [u[0] for u in uu], fs, [u[2] * fs for u in uu], [u[3] for u in uu], fft_len=fft_len, mag_dim=mag_dim,
                phase_dim=phase_dim, b_const_rate=b_const_rate,
                alpha_phase=False)     # the reference passes alpha_phase=b_mag_fbank_mel (= False, i.e. 0.0) at src/magphase.py:3010 -- replicated
            try:
                outs = call(utts)
            except (ValueError, IndexError):
                if on_error != 'skip':
                    raise
                # an utterance with unusable marks fails the argument checks of the whole batch: find it one by one
                g2, outs = [], []
                for t, u in zip(g, utts):
                    try:
                        outs.extend(call([u]))
                        g2.append(t)
                    except (ValueError, IndexError) as e:
                        crashes.add(t, e)
                g = g2
            for f in writes:
                f.result()                                                    # surface write errors of batch k-1
            writes = [pool.submit(_write_features, out_feats_dir, t, o, b_const_rate) for t, o in zip(g, outs)]
            n_frames += sum(o[0].shape[0] for o in outs)
            n_done += len(g)
            if verbose:
                print('analysed %d / %d utterances' % (n_done, len(tokens)))
        for f in writes:
            f.result()
    return dict(utterances=n_done, frames=n_frames, seconds=time.perf_counter() - t0, skipped=skipped,
                failed=crashes.tokens)


def _load_features(in_feats_dir, token, mag_dim, phase_dim):
    rd = lambda ext, dim: io.read_binfile(os.path.join(in_feats_dir, token + ext), dim=dim)
    return rd('.mag', mag_dim), rd('.real', phase_dim), rd('.imag', phase_dim), rd('.lf0', 1)


def _noise_draws(v_lf0, fs, fft_len, b_const_rate):
    """How many np.random.uniform draws synthesis_from_compressed takes for this utterance (src/magphase.py:879-883)."""
    v_lf0 = np.asarray(v_lf0, dtype=np.float64).reshape(-1)
    _, ns_lens = mp.compressed_synthesis_geometry([v_lf0], [v_lf0.size], fs, fft_len if fft_len else mp.define_fft_len(fs),
                                                  b_const_rate=b_const_rate)
    return int(ns_lens[0])


def run_waveform_generation(tokens, in_feats_dir, out_syn_dir, mag_dim, phase_dim, fs, fft_len=None, pf_type='magphase',
                            b_const_rate=False, batch_utts=64, io_threads=8, verbose=False, resume=False, on_error='raise'):
    """synthesis_from_acoustic_modelling (src/magphase.py:3229-3275) for a list of file tokens (or an .scp path).
    The aperiodic noise is drawn from NumPy's global stream in list order (seed it for reproducible output); with
    ``resume`` the draws of the tokens whose wav already exists are consumed all the same, so a resumed run writes the
    files of an uninterrupted one.  resume / on_error: module docstring.
    Returns {'utterances', 'frames', 'seconds', 'skipped', 'failed'}."""
    _check_on_error(on_error)
    if isinstance(tokens, str):
        tokens = read_tokens(tokens)
    tokens = list(tokens)
    if pf_type not in ('magphase', 'merlin', 'no'):
        raise ValueError("pf_type must be 'magphase', 'merlin' or 'no'")
    os.makedirs(out_syn_dir, exist_ok=True)
    have = set(t for t in tokens if resume and _complete([os.path.join(out_syn_dir, t + '.wav')]))
    crashes = _CrashList(out_syn_dir)
    t0 = time.perf_counter()
    n_frames = n_done = 0
    with cf.ThreadPoolExecutor(max_workers=io_threads) as pool:
        groups = list(_batches(tokens, batch_utts))
        load = lambda g: [pool.submit(_load_features, in_feats_dir, t, mag_dim, phase_dim) for t in g]
        pending = load(groups[0]) if groups else []
        writes = []
        for k, g in enumerate(groups):
            g, feats = _collect(pending, g, on_error, crashes)
            pending = load(groups[k + 1]) if k + 1 < len(groups) else []
            if pf_type in ('magphase', 'merlin') and feats:
                # both post-filters work frame by frame (src/magphase.py:2300-2378, 3375-3465): one call over the stacked rows
                rows = np.concatenate([np.atleast_2d(f[0]) for f in feats], axis=0)
                rows = mp.post_filter(rows, fs) if pf_type == 'magphase' else mp.post_filter_merlin(rows, fs)
                off = np.concatenate(([0], np.cumsum([np.atleast_2d(f[0]).shape[0] for f in feats])))
                feats = [(rows[off[i]:off[i + 1]],) + f[1:] for i, f in enumerate(feats)]
            # maximal stretches of tokens still to do, in list order; the finished ones in between only advance the stream
            i = 0
            while i < len(g):
                if g[i] in have:
                    np.random.uniform(-1, 1, _noise_draws(feats[i][3], fs, fft_len, b_const_rate))
                    i += 1
                    continue
                j = i
                while j < len(g) and g[j] not in have:
                    j += 1
                run_t, run_f = g[i:j], feats[i:j]
                i = j
                try:
                    ys = mp.synthesis_from_compressed_batch(run_f, fs, fft_len=fft_len, b_const_rate=b_const_rate)
                except (ValueError, IndexError):
                    if on_error != 'skip':
                        raise
                    # the argument checks run before any noise is drawn: retry utterance by utterance to find the bad one
                    ys, ok = [], []
                    for t, f in zip(run_t, run_f):
                        try:
                            ys.append(np.array(mp.synthesis_from_compressed_batch([f], fs, fft_len=fft_len,
                                                                                  b_const_rate=b_const_rate)[0]))
                            ok.append(t)
                        except (ValueError, IndexError) as e:
                            crashes.add(t, e)
                    run_t = ok
                for f in writes:
                    f.result()
                # (copies: the batch result is a view into a pooled page-locked block that the next batch reuses)
                writes = [pool.submit(_write_wav, os.path.join(out_syn_dir, t + '.wav'), np.array(y), fs)
                          for t, y in zip(run_t, ys)]
                n_done += len(run_t)
                n_frames += sum(np.atleast_2d(f[0]).shape[0] for t, f in zip(g, feats) if t in set(run_t))
            if verbose:
                print('synthesised %d / %d utterances' % (n_done, len(tokens) - len(have)))
        for f in writes:
            f.result()
    return dict(utterances=n_done, frames=n_frames, seconds=time.perf_counter() - t0,
                skipped=[t for t in tokens if t in have], failed=crashes.tokens)


def run_chain_stream(batches, fs, fft_len=None, mag_dim=60, phase_dim=45, b_out_hpf=False, n_workers=2, seed=0,
                     sig_as_pcm16=True, out_dtype=np.float32, keep_outputs=False, n_inflight=None):
    """In-memory streaming driver: analysis_compressed -> synthesis_from_compressed for a sequence of batches, each batch a
    list of (v_sig, v_pm_smpls, v_voi).  ``n_workers`` host threads take the batches round-robin; every worker owns a
    private library context on the GPU (``_lib.set_thread_slot``), so the NumPy bookkeeping, the PCIe copies and the kernels
    of different batches overlap -- what the reference gets from one forked process per utterance (src/libutils.py:32-63).
    The noise of batch k comes from ``np.random.RandomState(seed + k)`` (independent of which worker runs it).
    ``n_inflight`` (default: no limit) is an optional admission gate of the device calls (``_lib.set_device_gate``).
    Measured on a B200 (128 x 5 s utterances per batch, PCM16 in / float32 out): 1 / 2 / 3 / 4 / 6 workers give 8.5 / 14.5 /
    15.7 / 17.3 / 17.6 M frames/s; a gate below the worker count never helped (one call in flight keeps the GPU busy only
    about half of the time: the first group's upload and the last group's download are not covered).
    Returns {'utterances', 'frames', 'seconds'} (+ 'outputs': per batch (features, waveforms) when keep_outputs)."""
    from . import _lib
    batches = list(batches)
    results = [None] * len(batches)

    def work(w):
        _lib.set_thread_slot(w)
        outs = ys = None
        for k in range(w, len(batches), n_workers):
            b = batches[k]
            del outs, ys                 # the previous batch's page-locked result blocks go back to the pool first
            sigs = [u[0] for u in b]
            if sig_as_pcm16 and all(np.asarray(x).dtype != np.int16 for x in sigs):
                sigs = [np.round(np.asarray(x) * 32768.0).astype(np.int16) for x in sigs]
            outs = mp.analysis_compressed_batch(sigs, fs, [u[1] for u in b], [u[2] for u in b], fft_len=fft_len, mag_dim=mag_dim,
                                                phase_dim=phase_dim, out_dtype=out_dtype)
            ys = mp.synthesis_from_compressed_batch([o[:4] for o in outs], fs, fft_len=fft_len, b_out_hpf=b_out_hpf,
                                                    out_dtype=out_dtype, rng=np.random.RandomState(seed + k))
            frames = sum(o[4].size for o in outs)
            results[k] = (frames, ([tuple(np.array(a) for a in o[:5]) for o in outs], [np.array(y) for y in ys]) if keep_outputs else None)

    # page-lock the result arenas up front: every worker holds the features and the waveform of one batch (float32: 150
    # values per frame + one per sample), twice that for the order in which blocks of different sizes come and go
    peak = max((sum(4 * (np.size(u[0]) + 150 * np.size(u[1])) for u in b) for b in batches), default=0)
    _lib.pinned.ensure(2 * n_workers * peak * (np.dtype(out_dtype).itemsize // 4 or 1))
    prev_gate = _lib.set_device_gate(n_inflight if n_inflight and n_inflight < n_workers else None)
    t0 = time.perf_counter()
    try:
        with cf.ThreadPoolExecutor(max_workers=n_workers) as pool:
            for f in [pool.submit(work, w) for w in range(n_workers)]:
                f.result()
    finally:
        _lib.set_device_gate(prev_gate)
    r = dict(utterances=sum(len(b) for b in batches), frames=sum(x[0] for x in results), seconds=time.perf_counter() - t0)
    if keep_outputs:
        r['outputs'] = [x[1] for x in results]
    return r


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest='cmd', required=True)
    e = sub.add_parser('extract', help='wav (+ .est) -> .mag .real .imag .lf0 .shift')
    e.add_argument('--scp', required=True)
    e.add_argument('--wav-dir', required=True)
    e.add_argument('--out-dir', required=True)
    e.add_argument('--est-dir', default=None, help='REAPER .est files (<token>.est); default: run the REAPER binary')
    g = sub.add_parser('generate', help='.mag .real .imag .lf0 -> wav')
    g.add_argument('--scp', required=True)
    g.add_argument('--feats-dir', required=True)
    g.add_argument('--out-dir', required=True)
    g.add_argument('--fs', type=int, default=48000)
    g.add_argument('--pf-type', default='magphase', choices=['magphase', 'merlin', 'no'])
    g.add_argument('--seed', type=int, default=None, help='np.random.seed for the aperiodic noise')
    for p in (e, g):
        p.add_argument('--mag-dim', type=int, default=60)
        p.add_argument('--phase-dim', type=int, default=45)
        p.add_argument('--const-rate', action='store_true')
        p.add_argument('--batch-utts', type=int, default=64)
        p.add_argument('--io-threads', type=int, default=8)
        p.add_argument('--resume', action='store_true', help='skip tokens whose output files are already complete')
        p.add_argument('--skip-errors', action='store_true',
                       help='list failing tokens in crash_file_list_<host>_<pid>.scp and go on')
    a = ap.parse_args(argv)
    if a.cmd == 'extract':
        r = run_feature_extraction(a.scp, a.wav_dir, a.out_dir, est_dir=a.est_dir, mag_dim=a.mag_dim, phase_dim=a.phase_dim,
                                   b_const_rate=a.const_rate, batch_utts=a.batch_utts, io_threads=a.io_threads, verbose=True,
                                   resume=a.resume, on_error='skip' if a.skip_errors else 'raise')
    else:
        if a.seed is not None:
            np.random.seed(a.seed)
        r = run_waveform_generation(a.scp, a.feats_dir, a.out_dir, a.mag_dim, a.phase_dim, a.fs, pf_type=a.pf_type,
                                    b_const_rate=a.const_rate, batch_utts=a.batch_utts, io_threads=a.io_threads, verbose=True,
                                    resume=a.resume, on_error='skip' if a.skip_errors else 'raise')
    if r['skipped'] or r['failed']:
        print('%d tokens already done, %d failed' % (len(r['skipped']), len(r['failed'])))
    print('Done! %d utterances, %d frames in %.2f s (%.0f frames/s incl. file IO)'
          % (r['utterances'], r['frames'], r['seconds'], r['frames'] / max(r['seconds'], 1e-9)))


if __name__ == '__main__':
    main()

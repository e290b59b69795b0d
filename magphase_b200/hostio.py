"""File formats either side of the hot path (host only): PCM wav, raw float32 feature files, REAPER .est.

Mirrors src/libutils.py:112-127 (read_binfile / write_binfile), src/libaudio.py:343-365 (wav IO through
``soundfile``, which is not installed here -> ``scipy.io.wavfile``) and src/libaudio.py:421-447 (.est reader).
"""
import configparser
import os
import socket

import numpy as np


def read_audio_file(filepath):
    """float64 in [-1, 1) + sample rate, like ``soundfile.read`` on PCM data."""
    try:
        import soundfile as sf
        return sf.read(filepath)
    except ImportError:
        pass
    from scipy.io import wavfile
    fs, data = wavfile.read(filepath)
    if data.dtype == np.int16:
        data = data.astype(np.float64) / 32768.0
    elif data.dtype == np.int32:
        data = data.astype(np.float64) / 2147483648.0
    elif data.dtype == np.uint8:
        data = (data.astype(np.float64) - 128.0) / 128.0
    else:
        data = data.astype(np.float64)
    return data, fs


def write_audio_file(filepath, v_signal, fs, norm=0.98):
    """Peak-normalise to ``norm`` (None: no normalisation) and write PCM16.  src/libaudio.py:352-365"""
    v_signal = np.asarray(v_signal, dtype=np.float64)
    if norm is not None:
        v_signal = norm * v_signal / np.max(np.abs(v_signal))
    try:
        import soundfile as sf
        sf.write(filepath, v_signal, fs)
        return
    except ImportError:
        pass
    from scipy.io import wavfile
    pcm = np.clip(np.round(v_signal * 32768.0), -32768, 32767).astype(np.int16)
    wavfile.write(filepath, int(fs), pcm)


def read_binfile(filename, dim=60):
    """Raw float32, row-major, no header -> float64 (n, dim), squeezed.  src/libutils.py:112-120"""
    v_data = np.fromfile(filename, dtype=np.float32)
    if np.mod(v_data.size, dim) != 0:
        raise ValueError('Dimension provided not compatible with file size.')
    return np.squeeze(v_data.reshape((-1, dim)).astype('float64'))


def write_binfile(m_data, filename):
    """src/libutils.py:122-127"""
    np.array(m_data, 'float32').tofile(filename)


def read_reaper_est_file(est_file, check_len_smpls=-1, fs=-1, skiprows=7, usecols=(0, 1)):
    """Times (s) and voicing flags of a REAPER .est file; drops non-increasing times and marks at or
    beyond the last sample.  src/libaudio.py:421-447"""
    if (check_len_smpls > 0) and (fs == -1):
        raise ValueError('If check_len_smpls given, fs must be provided as well.')
    if skiprows == 7:
        # the reference skips a fixed 7 lines (REAPER's usual header); follow the header's own end marker when the
        # file carries one, so that extra header fields do not shift the data
        with open(est_file) as f:
            for i, line in enumerate(f):
                if line.strip() == 'EST_Header_End':
                    skiprows = i + 1
                    break
                if i > 64:
                    break
    m_data = np.loadtxt(est_file, skiprows=skiprows, usecols=list(usecols), ndmin=2)
    if m_data.size == 0:
        raise ValueError('%s holds no pitch marks' % est_file)
    v_pm_sec, v_voi = m_data[:, 0], m_data[:, 1]
    ok = np.hstack((True, np.diff(v_pm_sec) > 0))
    v_pm_sec, v_voi = v_pm_sec[ok], v_voi[ok]
    if check_len_smpls > 0:
        v_pm_smpls = np.round(v_pm_sec * fs).astype(int)
        if v_pm_smpls[-1] >= (check_len_smpls - 1):
            ok2 = v_pm_smpls < (check_len_smpls - 1)
            v_pm_sec, v_voi = v_pm_sec[ok2], v_voi[ok2]
    return v_pm_sec, v_voi


def write_reaper_est_file(est_file, v_pm_sec, v_voi):
    """Minimal writer of the same text layout (7 header lines), used by tests and the synthetic demos."""
    with open(est_file, 'w') as f:
        f.write('EST_File Track\nDataType ascii\nNumFrames %d\nNumChannels 1\nFrameShift 0.00000\n'
                'VoicingEnabled true\nEST_Header_End\n' % len(v_pm_sec))
        for t, v in zip(v_pm_sec, v_voi):
            f.write('%.6f %d\n' % (t, int(v)))


def find_tool(name):
    """Path of an external tool (REAPER): config.ini [TOOLS] bin_dir, else <repo>/tools/bin.  src/libaudio.py:20-34"""
    here = os.path.dirname(os.path.realpath(__file__))
    cand = [os.path.realpath(os.path.join(here, '..', 'tools', 'bin', name))]
    cfg = configparser.ConfigParser()
    if cfg.read(os.path.join(here, '..', 'config.ini')):
        try:
            bin_dir = cfg.get('TOOLS', 'bin_dir')
            if bin_dir != '':
                cand.insert(0, os.path.join(bin_dir, name))
        except (configparser.NoSectionError, configparser.NoOptionError):
            pass
    for c in cand:
        if os.path.isfile(c) and os.access(c, os.X_OK):
            return c
    return None


def ins_pid(filename):
    """temp-file name unique per host and process.  src/libutils.py:187-195"""
    base, ext = os.path.splitext(filename)
    return '%s_%s_%d%s' % (base, socket.gethostname(), os.getpid(), ext)

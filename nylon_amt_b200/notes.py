"""Note decoding: frame-level onset / offset / mpe / velocity arrays -> note list.

Host-side restructuring of AMT.mpe2note (reference hftt_code/model/amt.py:179-344): the reference scans every
(frame, pitch) cell in pure Python with inner left/right scans; here each pitch column is handled with run-length
arithmetic in numpy and only the (few) detected onsets are visited in Python.  Semantics kept exactly:
  * a frame is a peak if it is >= threshold and the first differing value on each side is lower (plateau points all
    count, amt.py:196-212)
  * sub-frame peak time from the two direct neighbours (amt.py:213-222), in the same float32 arithmetic numpy >= 2
    performs on float32 inputs
  * note end = offset peak after the onset, clipped to the next onset; or the first frame with mpe < threshold;
    'shorter' / 'longer' / 'offset' selection (amt.py:287-334); notes with velocity 0 dropped unless
    mode_velocity != 'ignore_zero'; overlapping same-pitch notes are cut (amt.py:336-339); sorted by onset, then pitch.
"""
import numpy as np


def _detect_peaks(col, thr, hop_sec):
    """Peaks of one pitch column -> (locs int array, times float32-valued python floats)."""
    n = col.shape[0]
    if n == 0:
        return np.zeros(0, np.int64), []
    # run-length encode equal values
    change = np.flatnonzero(col[1:] != col[:-1]) + 1
    run_start = np.concatenate(([0], change))
    run_val = col[run_start]
    run_id = np.zeros(n, np.int64)
    run_id[change] = 1
    run_id = np.cumsum(run_id)
    left_lower = np.ones(run_val.shape[0], bool)
    right_lower = np.ones(run_val.shape[0], bool)
    left_lower[1:] = run_val[1:] > run_val[:-1]
    right_lower[:-1] = run_val[:-1] > run_val[1:]
    is_peak_run = left_lower & right_lower
    cand = np.flatnonzero((col >= thr) & is_peak_run[run_id])
    times = []
    hop32 = np.float32(hop_sec * 0.5)
    for i in cand:
        t = np.float32(i * hop_sec)
        if 0 < i < n - 1:
            a, b, c = col[i - 1], col[i], col[i + 1]
            if a == c:
                pass
            elif a > c:
                t = np.float32(i * hop_sec) - np.float32(hop32 * np.float32(a - c) / np.float32(b - c))
                times.append(float(t))
                continue
            else:
                t = np.float32(i * hop_sec) + np.float32(hop32 * np.float32(c - a) / np.float32(b - a))
                times.append(float(t))
                continue
            times.append(i * hop_sec)
        else:
            times.append(i * hop_sec)
    return cand, times


def mpe2note(config, a_onset=None, a_offset=None, a_mpe=None, a_velocity=None, thred_onset=0.5, thred_offset=0.5, thred_mpe=0.5,
             mode_velocity='ignore_zero', mode_offset='shorter'):
    a_onset = np.asarray(a_onset)
    a_offset = np.asarray(a_offset)
    a_mpe = np.asarray(a_mpe)
    a_velocity = np.asarray(a_velocity)
    hop_sec = float(config['feature']['hop_sample'] / config['feature']['sr'])
    n_frames_mpe = a_mpe.shape[0]
    a_note = []
    for j in range(config['midi']['num_note']):
        on_loc, on_time = _detect_peaks(a_onset[:, j], thred_onset, hop_sec)
        if on_loc.size == 0:
            continue
        off_loc, off_time = _detect_peaks(a_offset[:, j], thred_offset, hop_sec)
        below = np.flatnonzero(a_mpe[:, j] < thred_mpe)
        for idx_on in range(on_loc.size):
            loc_onset = int(on_loc[idx_on])
            time_onset = on_time[idx_on]
            if idx_on + 1 < on_loc.size:
                loc_next = int(on_loc[idx_on + 1])
                time_next = on_time[idx_on + 1]
            else:
                loc_next = n_frames_mpe
                time_next = (loc_next - 1) * hop_sec
            # first offset peak strictly after the onset
            k = int(np.searchsorted(off_loc, loc_onset, side='right'))
            flag_offset = k < off_loc.size
            loc_offset, time_offset = loc_onset + 1, 0.0
            if flag_offset:
                loc_offset, time_offset = int(off_loc[k]), off_time[k]
            if loc_offset > loc_next:
                loc_offset, time_offset = loc_next, time_next
            # first frame in (loc_onset, loc_next) whose mpe falls below the threshold
            kb = int(np.searchsorted(below, loc_onset, side='right'))
            flag_mpe = kb < below.size and int(below[kb]) < loc_next
            loc_mpe, time_mpe = loc_onset + 1, 0.0
            if flag_mpe:
                loc_mpe = int(below[kb])
                time_mpe = loc_mpe * hop_sec
            pitch_value = int(j + config['midi']['note_min'])
            velocity_value = int(a_velocity[loc_onset][j])
            if (not flag_offset) and (not flag_mpe):
                offset_value = float(time_next)
            elif flag_offset and (not flag_mpe):
                offset_value = float(time_offset)
            elif (not flag_offset) and flag_mpe:
                offset_value = float(time_mpe)
            elif mode_offset == 'offset':
                offset_value = float(time_offset)
            elif mode_offset == 'longer':
                offset_value = float(time_offset) if loc_offset >= loc_mpe else float(time_mpe)
            else:
                offset_value = float(time_offset) if loc_offset <= loc_mpe else float(time_mpe)
            if mode_velocity != 'ignore_zero' or velocity_value > 0:
                a_note.append({'pitch': pitch_value, 'onset': float(time_onset), 'offset': offset_value, 'velocity': velocity_value})
            if (len(a_note) > 1) and (a_note[-1]['pitch'] == a_note[-2]['pitch']) and (a_note[-1]['onset'] < a_note[-2]['offset']):
                a_note[-2]['offset'] = a_note[-1]['onset']
    a_note = sorted(sorted(a_note, key=lambda x: x['pitch']), key=lambda x: x['onset'])
    return a_note


# ---- device-assisted variant ---------------------------------------------------------------------------------------------
def _peaks_device(a_dev, thr, hop_sec):
    """Peaks of every pitch column of a device map [T, N]: (pitch int64[n], loc int64[n], time float64[n]) sorted by (pitch, loc)."""
    import ctypes
    import torch
    from . import _lib
    L = _lib.lib()
    T, N = a_dev.shape
    flags = torch.empty((N, T), dtype=torch.uint8, device=a_dev.device)
    stream = torch.cuda.current_stream(a_dev.device).cuda_stream
    _lib.check(L.hft_note_peaks(ctypes.c_void_p(a_dev.data_ptr()), T, N, float(thr), ctypes.c_void_p(flags.data_ptr()), ctypes.c_void_p(stream)), "hft_note_peaks")
    idx = torch.nonzero(flags).contiguous()                    # [n, 2] int64, row-major order = sorted by (pitch, frame)
    n = idx.shape[0]
    kind = torch.empty(n, dtype=torch.uint8, device=a_dev.device)
    t32 = torch.empty(n, dtype=torch.float32, device=a_dev.device)
    _lib.check(L.hft_note_peak_times(ctypes.c_void_p(a_dev.data_ptr()), T, N, ctypes.c_void_p(idx.data_ptr()), n, float(hop_sec),
                                     ctypes.c_void_p(kind.data_ptr()), ctypes.c_void_p(t32.data_ptr()), ctypes.c_void_p(stream)), "hft_note_peak_times")
    idx_h = idx.cpu().numpy()
    pitch, loc = idx_h[:, 0], idx_h[:, 1]
    time = np.where(kind.cpu().numpy() == 1, t32.cpu().numpy().astype(np.float64), loc * hop_sec)
    return idx, pitch, loc, time


def mpe2note_device(config, a_onset=None, a_offset=None, a_mpe=None, a_velocity=None, thred_onset=0.5, thred_offset=0.5, thred_mpe=0.5,
                    mode_velocity='ignore_zero', mode_offset='shorter', device='cuda'):
    """Same result as mpe2note (and as the reference, amt.py:179-344), with the scans over the [T, n_note] maps done by
    libhft_sm100 (hft_note_peaks / hft_note_peak_times / hft_note_first_below) and the note assembly vectorised over the detected onsets."""
    import ctypes
    import torch
    from . import _lib
    hop_sec = float(config['feature']['hop_sample'] / config['feature']['sr'])
    note_min = config['midi']['note_min']
    dev = torch.device(device)
    on = torch.as_tensor(np.ascontiguousarray(a_onset, dtype=np.float32)).to(dev)
    off = torch.as_tensor(np.ascontiguousarray(a_offset, dtype=np.float32)).to(dev)
    mpe = torch.as_tensor(np.ascontiguousarray(a_mpe, dtype=np.float32)).to(dev)
    a_velocity = np.asarray(a_velocity)
    T, N = on.shape
    n_frames_mpe = mpe.shape[0]
    with torch.cuda.device(dev):
        idx_on, p_on, l_on, t_on = _peaks_device(on, thred_onset, hop_sec)
        _, p_off, l_off, t_off = _peaks_device(off, thred_offset, hop_sec)
        m = p_on.shape[0]
        if m == 0:
            return []
        # next onset of the same pitch (amt.py:262-268)
        has_next = np.zeros(m, bool)
        has_next[:-1] = p_on[1:] == p_on[:-1]
        loc_next = np.where(has_next, np.roll(l_on, -1), n_frames_mpe)
        time_next = np.where(has_next, np.roll(t_on, -1), (n_frames_mpe - 1) * hop_sec)
        limit = torch.from_numpy(np.ascontiguousarray(loc_next.astype(np.int64))).to(dev)
        below = torch.empty(m, dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().hft_note_first_below(ctypes.c_void_p(mpe.data_ptr()), n_frames_mpe, N, ctypes.c_void_p(idx_on.data_ptr()), ctypes.c_void_p(limit.data_ptr()),
                                                   m, float(thred_mpe), ctypes.c_void_p(below.data_ptr()), ctypes.c_void_p(stream)), "hft_note_first_below")
        below = below.cpu().numpy()
    # first offset peak strictly after the onset, same pitch (amt.py:274-283): search on the combined (pitch, frame) key
    stride = max(T, n_frames_mpe) + 2
    key_off = p_off * stride + l_off
    k = np.searchsorted(key_off, p_on * stride + l_on, side='right')
    kc = np.minimum(k, max(key_off.shape[0] - 1, 0))
    flag_offset = (k < key_off.shape[0]) & ((p_off[kc] if key_off.shape[0] else np.zeros(m, np.int64)) == p_on) if key_off.shape[0] else np.zeros(m, bool)
    loc_offset = np.where(flag_offset, l_off[kc] if key_off.shape[0] else 0, l_on + 1)
    time_offset = np.where(flag_offset, t_off[kc] if key_off.shape[0] else 0.0, 0.0)
    clip = loc_offset > loc_next
    loc_offset = np.where(clip, loc_next, loc_offset)
    time_offset = np.where(clip, time_next, time_offset)
    flag_mpe = below >= 0
    loc_mpe = np.where(flag_mpe, below, l_on + 1)
    time_mpe = np.where(flag_mpe, loc_mpe * hop_sec, 0.0)
    if mode_offset == 'offset':
        both = time_offset
    elif mode_offset == 'longer':
        both = np.where(loc_offset >= loc_mpe, time_offset, time_mpe)
    else:
        both = np.where(loc_offset <= loc_mpe, time_offset, time_mpe)
    offset_value = np.where(~flag_offset & ~flag_mpe, time_next, np.where(flag_offset & ~flag_mpe, time_offset, np.where(~flag_offset & flag_mpe, time_mpe, both)))
    velocity = a_velocity[l_on, p_on].astype(np.int64)
    keep = np.ones(m, bool) if mode_velocity != 'ignore_zero' else velocity > 0
    p_k, on_k, off_k, v_k = p_on[keep], t_on[keep].astype(np.float64), offset_value[keep].astype(np.float64), velocity[keep]
    # overlapping notes of the same pitch are cut at the next (kept) onset (amt.py:336-339)
    if p_k.shape[0] > 1:
        same = p_k[1:] == p_k[:-1]
        cut = same & (on_k[1:] < off_k[:-1])
        off_k[:-1] = np.where(cut, on_k[1:], off_k[:-1])
    order = np.argsort(on_k, kind='stable')                     # sorted by pitch already; stable sort by onset = the reference's double sort
    return [{'pitch': int(p_k[i] + note_min), 'onset': float(on_k[i]), 'offset': float(off_k[i]), 'velocity': int(v_k[i])} for i in order]

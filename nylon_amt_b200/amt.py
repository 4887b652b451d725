"""Host-side mirror of the reference's hftt_code/model/amt.py (class AMT) for the hot path.

Same constructor and method signatures as the reference (AMT.__init__ amt.py:10, wav2feature :34, transcript :66,
transcript_stride :121, mpe2note :179).  The arithmetic runs in libhft_sm100.so (include/hft_sm100.h):
  * wav2feature  -> fused log-mel kernel (replaces torchaudio MelSpectrogram + log, amt.py:59-61)
  * transcript   -> windows are gathered on the device as a strided view of the padded feature and run through
                    Model_SPEC2MIDI.forward in batches (replaces the batch-1 loop with 1 H2D + 8 D2H per segment,
                    amt.py:88-113)
There is no CPU fallback: without a CUDA device these methods raise.
"""
import ctypes
import pickle
import wave as _wave

import numpy as np
import torch

from . import _lib, melfb


def _read_wav(path):
    """(float32 [C,N] in [-1,1), sample_rate) like torchaudio.load's default normalisation (amt.py:55)."""
    try:
        from scipy.io import wavfile
        sr, data = wavfile.read(path)
        if data.ndim == 1:
            data = data[:, None]
        if data.dtype == np.int16:
            x = data.astype(np.float32) / np.float32(32768.0)
        elif data.dtype == np.int32:
            x = (data.astype(np.float64) / 2147483648.0).astype(np.float32)
        elif data.dtype == np.uint8:
            x = (data.astype(np.float32) - np.float32(128.0)) / np.float32(128.0)
        else:
            x = data.astype(np.float32)
        return torch.from_numpy(np.ascontiguousarray(x.T)), int(sr)
    except ImportError:
        with _wave.open(path, "rb") as f:
            n_ch, width, sr, n = f.getnchannels(), f.getsampwidth(), f.getframerate(), f.getnframes()
            raw = f.readframes(n)
        if width != 2:
            raise RuntimeError("only 16-bit PCM wav is readable without scipy")
        a = np.frombuffer(raw, dtype="<i2").reshape(-1, n_ch).T.astype(np.float32) / np.float32(32768.0)
        return torch.from_numpy(np.ascontiguousarray(a)), int(sr)


class _LogmelPlan:
    def __init__(self, window, fb, log_offset):
        self.ptr = ctypes.c_void_p()
        w = window.contiguous().cpu().float()
        f = fb.contiguous().cpu().float()
        _lib.check(_lib.lib().hft_logmel_create(ctypes.byref(self.ptr), ctypes.c_void_p(w.data_ptr()), ctypes.c_void_p(f.data_ptr()),
                                                ctypes.c_float(log_offset)), "hft_logmel_create")

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().hft_logmel_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class _ResamplePlan:
    def __init__(self, orig_hz, new_hz):
        self.ptr = ctypes.c_void_p()
        _lib.check(_lib.lib().hft_resample_create_hz(ctypes.byref(self.ptr), int(orig_hz), int(new_hz)), "hft_resample_create_hz")

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().hft_resample_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class AMT():
    def __init__(self, config, model_path, batch_size=1, verbose_flag=False):
        if verbose_flag is True:
            print('torch version: ' + torch.__version__)
            print('torch cuda   : ' + str(torch.cuda.is_available()))
        # the reference falls back to 'cpu' (amt.py:14-17); this implementation is CUDA-only and fails at first use
        self.device = 'cuda'
        self.config = config
        if model_path is None:
            self.model = None
        else:
            from . import install_reference_aliases
            install_reference_aliases()
            with open(model_path, 'rb') as f:
                self.model = pickle.load(f)
            self.model = self.model.to(self.device)
            self.model.eval()
            if verbose_flag is True:
                print(self.model)
        self.batch_size = batch_size
        self._plan = None
        self._fb = None

    # ---- feature -----------------------------------------------------------------------------------------
    def _check_feature_config(self):
        f = self.config['feature']
        if not (f['sr'] == 16000 and f['fft_bins'] == 2048 and f['window_length'] == 2048 and f['hop_sample'] == 256 and
                f['mel_bins'] == 256 and f['pad_mode'] == 'constant'):
            raise RuntimeError("the fused log-mel kernel implements the reference geometry only "
                               "(sr 16000, n_fft/win 2048, hop 256, 256 mels, constant padding); got %r" % (f,))

    def mel_fb(self):
        if self._fb is None:
            f = self.config['feature']
            self._fb = melfb.melscale_fbanks(f['fft_bins'] // 2 + 1, 0.0, float(f['sr'] // 2), f['mel_bins'], f['sr'])
        return self._fb

    def _logmel_plan(self):
        if self._plan is None:
            self._check_feature_config()
            if not torch.cuda.is_available():
                raise RuntimeError("no CUDA device: the B200 path has no CPU fallback")
            self._plan = _LogmelPlan(melfb.hann_window(self.config['feature']['fft_bins']), self.mel_fb(),
                                     float(self.config['feature']['log_offset']))
        return self._plan

    def wave2feature(self, wave_mono_16k):
        """Device-resident variant: fp32 CUDA tensor [N] (mono, 16 kHz) -> CUDA tensor [T,256]."""
        plan = self._logmel_plan()
        x = wave_mono_16k
        if not x.is_cuda:
            raise RuntimeError("wave2feature expects a CUDA tensor (no CPU fallback); use wav2feature for files")
        x = x.contiguous().float()
        n = x.numel()
        T = _lib.lib().hft_logmel_num_frames(n)
        out = torch.empty((T, self.config['feature']['mel_bins']), device=x.device, dtype=torch.float32)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().hft_logmel_f32(plan.ptr, ctypes.c_void_p(x.data_ptr()), n, ctypes.c_void_p(out.data_ptr()), T,
                                                 ctypes.c_void_p(stream)), "hft_logmel_f32")
        return out

    def waves2features(self, waves):
        """Ragged batch in one launch (the file loop of conv_wav2fe.py:41-48): list of CUDA [N_i] -> list of [T_i,256]."""
        plan = self._logmel_plan()
        if len(waves) == 0:
            return []
        dev = waves[0].device
        lens = [int(w.numel()) for w in waves]
        starts, pos = [], 0
        for n in lens:                       # clip starts padded to 4 samples: every clip stays TMA-eligible
            starts.append(pos)
            pos += (n + 3) // 4 * 4
        flat = torch.zeros(max(pos, 1), device=dev, dtype=torch.float32)
        for w, s, n in zip(waves, starts, lens):
            flat[s:s + n] = w.reshape(-1).float()
        Ts = [int(_lib.lib().hft_logmel_num_frames(n)) for n in lens]
        out = torch.empty((sum(Ts), self.config['feature']['mel_bins']), device=dev, dtype=torch.float32)
        c_start = (ctypes.c_int64 * len(lens))(*starts)
        c_len = (ctypes.c_int64 * len(lens))(*lens)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().hft_logmel_batch_f32(plan.ptr, ctypes.c_void_p(flat.data_ptr()), c_start, c_len, len(lens),
                                                       ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(stream)), "hft_logmel_batch_f32")
        outs, r = [], 0
        for T in Ts:
            outs.append(out[r:r + T])
            r += T
        return outs

    def _resample_plan(self, sr):
        """Plan of torchaudio.transforms.Resample(sr, 16000) (amt.py:57): the polyphase table is built inside the library
        (hft_resample_create_hz; within 1e-7 of torchaudio's own table, tests/test_resample.py)."""
        plans = self.__dict__.setdefault("_rs_plans", {})
        if sr not in plans:
            plans[sr] = _ResamplePlan(int(sr), int(self.config['feature']['sr']))
        return plans[sr]

    def wave2mono16k(self, wave, sr):
        """amt.py:56-58 on the device: CUDA tensor [C,N] at `sr` Hz -> CUDA tensor [N'] mono at config sr (channel mean + sinc resampling)."""
        if not wave.is_cuda:
            raise RuntimeError("wave2mono16k expects a CUDA tensor (no CPU fallback)")
        w = wave.float().contiguous()
        if w.dim() == 1:
            w = w[None]
        if sr == self.config['feature']['sr']:
            return torch.mean(w, dim=0)                      # Resample.forward returns its input when the rates agree
        plan = self._resample_plan(sr)
        C, n = w.shape
        n_out = int(_lib.lib().hft_resample_num_samples(plan.ptr, n))
        out = torch.empty(n_out, device=w.device, dtype=torch.float32)
        stream = torch.cuda.current_stream(w.device).cuda_stream
        with torch.cuda.device(w.device):
            _lib.check(_lib.lib().hft_resample_mono_f32(plan.ptr, ctypes.c_void_p(w.data_ptr()), C, n, ctypes.c_void_p(out.data_ptr()), n_out,
                                                        ctypes.c_void_p(stream)), "hft_resample_mono_f32")
        return out

    def wav2feature(self, f_wav):
        """amt.py:34-63: wav file -> CPU FloatTensor [T, mel_bins] = log(mel + log_offset).T"""
        wave, sr = _read_wav(f_wav)
        if sr != self.config['feature']['sr']:
            # amt.py:56-58: channel mean + torchaudio Resample, on the device, feeding the log-mel kernel without a host round trip
            if not torch.cuda.is_available():
                raise RuntimeError("no CUDA device: the B200 path has no CPU fallback")
            return self.wave2feature(self.wave2mono16k(wave.cuda(), sr)).cpu()
        wave_mono = torch.mean(wave, dim=0)
        plan = self._logmel_plan()
        x = wave_mono.contiguous().float()
        n = x.numel()
        T = _lib.lib().hft_logmel_num_frames(n)
        out = torch.empty((T, self.config['feature']['mel_bins']), dtype=torch.float32)
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().hft_logmel_host_f32(plan.ptr, ctypes.c_void_p(x.data_ptr()), n, ctypes.c_void_p(out.data_ptr()), T,
                                                  ctypes.c_void_p(stream)), "hft_logmel_host_f32")
        return out

    # ---- transcription -----------------------------------------------------------------------------------
    def _run_windows(self, a_input, n_win, hop, n_offset, n_keep, n_out_rows, mode, ablation_flag):
        """Run the model over windows [i*hop, i*hop+W) of the padded device feature and keep rows
        [n_offset, n_offset+n_keep) of every window's outputs (amt.py:88-113 / :146-171)."""
        if self.model is None:
            raise RuntimeError("AMT was constructed without a model")
        if mode != 'combination' or ablation_flag:
            raise NotImplementedError("only mode='combination' of Model_SPEC2MIDI is on the B200 hot path")
        cfg = self.config
        n_note, n_bin = cfg['midi']['num_note'], cfg['feature']['n_bins']
        W = cfg['input']['margin_b'] + cfg['input']['num_frame'] + cfg['input']['margin_f']
        F = cfg['input']['num_frame']
        dev = a_input.device
        # segments per forward call: AMT.batch_size when given (the reference stores it and never reads it, amt.py:31), else 48
        # (measured on B200: 16 -> 2 210 x, 30 -> 2 255 x, 45..128 -> 2 260 x real-time; fewer kernel tails per hour of audio)
        chunk = int(self.batch_size) if (self.batch_size is not None and int(self.batch_size) > 1) else 48
        self.model.eval()
        res_f = [np.zeros((n_out_rows, n_note), dtype=np.float32) for _ in range(6)]
        res_v = [np.zeros((n_out_rows, n_note), dtype=np.int8) for _ in range(2)]
        spec_all = torch.as_strided(a_input, (n_win, n_bin, W), (hop * n_bin, 1, n_bin))
        nb = min(chunk, n_win)
        if (getattr(self.model, 'max_batch', None) or 0) < nb:
            self.model.max_batch = nb                       # one internal pass per call (the workspace is sized by it)
        # Device outputs and pinned host staging are double-buffered per slot: the D2H copies of chunk i run behind the forward of
        # chunk i+1, and the host only waits on the event of the slot it is about to reuse (the reference syncs 8 times per 2 s
        # of audio, amt.py:104-113).  The velocity logits themselves ([nb, F, n_note, V] x 2) are never materialised: only their
        # argmax leaves the heads GEMM.
        slots = self._result_slots(dev, nb, F, n_keep, n_note)
        sl = slice(n_offset, n_offset + n_keep)

        def drain(slot):
            r0, b = slot["r0"], slot["b"]
            slot["event"].synchronize()
            n = max(0, min(b * n_keep, n_out_rows - r0))
            for dst, src in zip(res_f + res_v, slot["host"]):
                dst[r0:r0 + n] = src[:n].numpy()
            slot["b"] = 0

        with torch.no_grad():
            for i, w0 in enumerate(range(0, n_win, chunk)):
                b = min(chunk, n_win - w0)
                slot = slots[i % len(slots)]
                if slot["b"]:
                    drain(slot)
                outs = [t[:b] if t is not None else None for t in slot["dev"]]
                vel = [t[:b] for t in slot["vel"]]
                self.model.forward_into(spec_all[w0:w0 + b], outs, want_attention=False, velocity_argmax=vel)
                srcs = [outs[k] for k in (0, 1, 2, 5, 6, 7)] + vel
                for dst, src in zip(slot["host"], srcs):
                    dst[:b * n_keep].view(b, n_keep, n_note).copy_(src[:, sl], non_blocking=True)
                slot["event"].record(torch.cuda.current_stream(dev))
                slot["r0"], slot["b"] = w0 * n_keep, b
            order = sorted((s_ for s_ in slots if s_["b"]), key=lambda s_: s_["r0"])
            for slot in order:
                drain(slot)
        return res_f[0], res_f[1], res_f[2], res_v[0], res_f[3], res_f[4], res_f[5], res_v[1]

    def _result_slots(self, dev, nb, F, n_keep, n_note, n_slots=3):
        """Per-chunk device outputs + pinned host staging + event, cached across calls (pinned allocation is slow)."""
        cache = self.__dict__.setdefault("_slots", {})
        key = (str(dev), nb, F, n_keep, n_note)
        if key not in cache:
            cache.clear()                                   # one geometry at a time: do not hoard pinned memory
            slots = []
            for _ in range(n_slots):
                f32 = lambda: torch.empty((nb, F, n_note), device=dev, dtype=torch.float32)
                slots.append(dict(
                    dev=[f32(), f32(), f32(), None, None, f32(), f32(), f32(), None],
                    vel=[torch.empty((nb, F, n_note), device=dev, dtype=torch.int8) for _ in range(2)],
                    host=[torch.empty((nb * n_keep, n_note), dtype=torch.float32).pin_memory() for _ in range(6)] +
                         [torch.empty((nb * n_keep, n_note), dtype=torch.int8).pin_memory() for _ in range(2)],
                    event=torch.cuda.Event(), b=0, r0=0))
            cache[key] = slots
        for s_ in cache[key]:
            s_["b"] = 0
        return cache[key]

    def _to_device_feature(self, a_feature):
        if isinstance(a_feature, torch.Tensor):
            t = a_feature.detach().float()
        else:
            t = torch.from_numpy(np.array(a_feature, dtype=np.float32))
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: the B200 path has no CPU fallback")
        return t.to(self.device)

    def transcript(self, a_feature, mode='combination', ablation_flag=False):
        """amt.py:66-118.  a_feature: [num_frame, n_mels] -> 8 numpy arrays [T+len_s, num_note]."""
        cfg = self.config
        feat = self._to_device_feature(a_feature)
        T, F = feat.shape[0], cfg['input']['num_frame']
        mb, mf = cfg['input']['margin_b'], cfg['input']['margin_f']
        len_s = int(np.ceil(T / F) * F) - T
        a_input = torch.full((mb + T + len_s + mf, cfg['feature']['n_bins']), float(cfg['input']['min_value']), device=feat.device,
                             dtype=torch.float32)
        a_input[mb:mb + T] = feat
        n_win = (T + F - 1) // F
        return self._run_windows(a_input, n_win, F, 0, F, T + len_s, mode, ablation_flag)

    def transcript_stride(self, a_feature, n_offset, mode='combination', ablation_flag=False):
        """amt.py:121-176: half-frame stride, keeps rows [n_offset, n_offset+64) of every window."""
        cfg = self.config
        feat = self._to_device_feature(a_feature)
        T, F = feat.shape[0], cfg['input']['num_frame']
        mb, mf = cfg['input']['margin_b'], cfg['input']['margin_f']
        half = int(F / 2)
        tmp_len = T + mb + mf + half
        len_s = int(np.ceil(tmp_len / half) * half) - tmp_len
        rows = (mb + n_offset) + T + (len_s + mf + (half - n_offset))
        a_input = torch.full((rows, cfg['feature']['n_bins']), float(cfg['input']['min_value']), device=feat.device, dtype=torch.float32)
        a_input[mb + n_offset: mb + n_offset + T] = feat
        n_win = (T + half - 1) // half
        return self._run_windows(a_input, n_win, half, n_offset, half, T + len_s, mode, ablation_flag)

    # ---- note decoding (host, amt.py:179-344) ---------------------------------------------------------------
    def mpe2note(self, a_onset=None, a_offset=None, a_mpe=None, a_velocity=None, thred_onset=0.5, thred_offset=0.5, thred_mpe=0.5,
                 mode_velocity='ignore_zero', mode_offset='shorter'):
        from . import notes
        # long transcripts: the O(T x 88) scans run on the device (hft_note_*), the result is identical to the host restructuring
        n_cells = int(np.asarray(a_onset).shape[0]) * int(np.asarray(a_onset).shape[1]) if a_onset is not None else 0
        if n_cells >= (1 << 20) and torch.cuda.is_available():
            return notes.mpe2note_device(self.config, a_onset, a_offset, a_mpe, a_velocity, thred_onset, thred_offset, thred_mpe, mode_velocity, mode_offset)
        return notes.mpe2note(self.config, a_onset, a_offset, a_mpe, a_velocity, thred_onset, thred_offset, thred_mpe, mode_velocity,
                              mode_offset)

    def note2midi(self, a_note, f_midi):
        """amt.py:347-355 (file output; needs pretty_midi like the reference)."""
        import pretty_midi
        midi = pretty_midi.PrettyMIDI()
        instrument = pretty_midi.Instrument(program=0)
        for note in a_note:
            instrument.notes.append(pretty_midi.Note(velocity=note['velocity'], pitch=note['pitch'], start=note['onset'], end=note['offset']))
        midi.instruments.append(instrument)
        midi.write(f_midi)
        return

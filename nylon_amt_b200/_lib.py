"""ctypes binding of libhft_sm100.so (include/hft_sm100.h).  No fallback: if the library is missing or a call
fails, a RuntimeError is raised."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HFT_LIB_PATH") or os.path.join(_HERE, "lib", "libhft_sm100.so")   # HFT_LIB_PATH: kernel-variant experiments only

PREC = {"fp32": 0, "bf16": 1, "fp16": 2, "fp16x3": 3, "mixed": 4}

c_float_p = ctypes.c_void_p


class hft_dims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("n_margin", "n_frame", "n_bin", "cnn_channel", "cnn_kernel", "hid_dim", "pf_dim",
                                               "n_enc_layers", "n_dec_layers", "n_heads", "n_note", "n_velocity")]


class hft_outputs(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("onset_A", "offset_A", "mpe_A", "velocity_A", "attention",
                                                "onset_B", "offset_B", "mpe_B", "velocity_B", "velocity_A_argmax", "velocity_B_argmax")]


# every symbol include/hft_sm100.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "hft_version": (ctypes.c_int, []),
    "hft_last_error": (ctypes.c_char_p, []),
    "hft_device_sm_count": (ctypes.c_int, []),
    "hft_logmel_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float]),
    "hft_logmel_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "hft_logmel_num_frames": (ctypes.c_int64, [ctypes.c_int64]),
    "hft_logmel_f32": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "hft_logmel_batch_f32": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64),
                                          ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "hft_logmel_host_f32": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "hft_resample_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]),
    "hft_resample_build_table": (ctypes.c_int64, [ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int32),
                                                  ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
    "hft_resample_create_hz": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int32, ctypes.c_int32]),
    "hft_resample_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "hft_resample_num_samples": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_int64]),
    "hft_resample_mono_f32": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "hft_note_peaks": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]),
    "hft_note_peak_times": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p]),
    "hft_note_first_below": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_float,
                                            ctypes.c_void_p, ctypes.c_void_p]),
    "hft_model_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(hft_dims)]),
    "hft_model_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "hft_model_num_weights": (ctypes.c_int, [ctypes.c_void_p]),
    "hft_model_weight_name": (ctypes.c_char_p, [ctypes.c_void_p, ctypes.c_int]),
    "hft_model_weight_numel": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_int]),
    "hft_model_set_weights": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_void_p]),
    "hft_model_set_max_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    "hft_model_release_workspace": (ctypes.c_int, [ctypes.c_void_p]),
    "hft_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                   ctypes.c_int32, ctypes.POINTER(hft_outputs), ctypes.c_void_p]),
    "hft_last_launch_count": (ctypes.c_int64, []),
    "hft_tc_linear": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                     ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "hft_tc_ffn": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "hft_tc_attention": (ctypes.c_int, [ctypes.c_int, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p]),
    "hft_model_param_floats": (ctypes.c_int64, [ctypes.c_void_p]),
    "hft_model_param_offset": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_int]),
    "hft_model_params": (ctypes.c_void_p, [ctypes.c_void_p]),
    "hft_model_refresh": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "hft_model_get_params": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "hft_trainer_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_int32]),
    "hft_trainer_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "hft_trainer_set_dropout": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_float, ctypes.c_uint32]),
    "hft_dropout_mask": (ctypes.c_int, [ctypes.c_float, ctypes.c_uint32, ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "hft_forward_encoder": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]),
    "hft_forward_decoder": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(hft_outputs), ctypes.c_void_p]),
    "hft_train_forward_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_void_p,
                                                  ctypes.c_void_p, ctypes.c_void_p]),
    "hft_train_forward_backward_n": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_void_p,
                                                    ctypes.c_void_p, ctypes.c_void_p]),
    "hft_train_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                         ctypes.POINTER(hft_outputs), ctypes.c_void_p]),
    "hft_train_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(hft_outputs),
                                          ctypes.c_void_p, ctypes.c_void_p]),
    "hft_train_attention": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_float, ctypes.c_uint32,
                                           ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]),
    "hft_train_linear": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                        ctypes.c_void_p, ctypes.c_int32, ctypes.c_float, ctypes.c_void_p]),
    "hft_train_linear_wgrad": (ctypes.c_int, [ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32,
                                              ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]),
    "hft_adam_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_float, ctypes.c_float,
                                     ctypes.c_float, ctypes.c_float, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p]),
    "hft_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "hft_probe_fp32_fma": (ctypes.c_int, [ctypes.c_int32, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.c_void_p]),
    "hft_profile_read": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]),
}

_lib = None


def lib():
    """dlopen the C-ABI library and bind every declared symbol (raises if the library or a symbol is missing)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError("libhft_sm100.so is not built (%s); run `python __graft_entry__.py` -- there is no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().hft_last_error()
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else ""))


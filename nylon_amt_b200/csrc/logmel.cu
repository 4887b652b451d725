// Fused log-mel kernel for sm_100a: frames 16 kHz audio (hop 256, n_fft 2048, zero "center" padding), applies
// the Hann window, runs the 2048-point real FFT in shared memory / registers, |X|^2, banded 256-bin mel sums
// and log(+offset), writing [T,256] fp32.  One launch replaces torchaudio.transforms.MelSpectrogram + torch.log
// in AMT.wav2feature (reference hftt_code/model/amt.py:59-61; parameters hftt_code/corpus/config.json:2-12).
//
// Data movement: a persistent CTA per SM walks blocks of kFramesPerBlock consecutive frames.  The block's
// overlapping sample span ((F-1)*256 + 2048 floats) is staged ONCE into shared memory by the TMA engine
// (cp.async.bulk, double buffered behind an mbarrier) so each sample is read from HBM once although it is used
// by 8 frames; the output row of a frame is written with coalesced 128-byte stores.  Algorithmic traffic:
// 1 KB in + 1 KB out per frame.
#include "common.cuh"
#include "logmel_core.cuh"
#include "hft_internal.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace hft {

constexpr int kWarps = 16;          // 16 warps / SM: the kernel is issue- and latency-bound, not bandwidth-bound
constexpr int kFramesPerBlock = 16;
constexpr int kSpan = (kFramesPerBlock - 1) * kHop + kNfft;   // floats staged per block

struct LogmelClip {
  const float* wav;     // device, clip start
  long long n_samples;
  float* out;           // device, [n_frames,256]
  long long n_frames;
};

struct LogmelParams {
  const float* window;
  const float2* tw2;
  const float2* twr;
  const float* melw;
  const uint32_t* melinfo;
  float log_offset;
  int n_clips;
  int n_blocks;
  const LogmelClip* clips;   // device array when n_clips > 1
  const int* blk_prefix;     // device [n_clips+1] when n_clips > 1
  LogmelClip single;         // used when n_clips == 1 (no descriptor upload on the hot path)
};

struct SmemLayout {
  static constexpr int stage = 0;                                            // 2 * kSpan floats
  static constexpr int win = stage + 2 * kSpan * 4;
  static constexpr int tw2 = win + kNfft * 4;
  static constexpr int twr = tw2 + 1024 * 8;
  static constexpr int melw = twr + 520 * 8;
  static constexpr int melinfo = melw + kMaxMelW * 4;
  static constexpr int warp0 = melinfo + kMelInfo * 4;
  static constexpr int tile_bytes = 32 * kTStride * 8;                       // T / Z; P[0..1024] overlays it once Z is consumed
  static constexpr int per_warp = tile_bytes;
  static constexpr int bars = warp0 + kWarps * per_warp;
  static constexpr int total = bars + 16;
};

struct BlockInfo {
  LogmelClip clip;
  long long t0;       // first frame of the block inside the clip
  int nf;             // frames in this block
  long long s0;       // first staged sample (may be negative)
  int span;           // floats staged
  bool bulk;          // TMA-eligible: fully inside the clip and 16-byte aligned
};

__device__ __forceinline__ BlockInfo resolve_block(const LogmelParams& p, int blk) {
  BlockInfo b;
  int local = blk;
  if (p.n_clips == 1) {
    b.clip = p.single;
  } else {
    int lo = 0, hi = p.n_clips;             // largest c with blk_prefix[c] <= blk
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (p.blk_prefix[mid] <= blk) lo = mid; else hi = mid;
    }
    b.clip = p.clips[lo];
    local = blk - p.blk_prefix[lo];
  }
  b.t0 = (long long)local * kFramesPerBlock;
  long long rem = b.clip.n_frames - b.t0;
  b.nf = rem < kFramesPerBlock ? (int)rem : kFramesPerBlock;
  b.s0 = b.t0 * kHop - kNfft / 2;
  b.span = (b.nf - 1) * kHop + kNfft;
  b.bulk = (b.s0 >= 0) && (b.s0 + b.span <= b.clip.n_samples) && ((reinterpret_cast<uintptr_t>(b.clip.wav + b.s0) & 15) == 0);
  return b;
}

__global__ void __launch_bounds__(kWarps * 32, 1) logmel_kernel(const __grid_constant__ LogmelParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* stage = reinterpret_cast<float*>(smem + SmemLayout::stage);
  float* s_win = reinterpret_cast<float*>(smem + SmemLayout::win);
  float2* s_tw2 = reinterpret_cast<float2*>(smem + SmemLayout::tw2);
  float2* s_twr = reinterpret_cast<float2*>(smem + SmemLayout::twr);
  float* s_melw = reinterpret_cast<float*>(smem + SmemLayout::melw);
  uint32_t* s_melinfo = reinterpret_cast<uint32_t*>(smem + SmemLayout::melinfo);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SmemLayout::bars);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float2* s_T = reinterpret_cast<float2*>(smem + SmemLayout::warp0 + warp * SmemLayout::per_warp);
  float* s_P = reinterpret_cast<float*>(s_T);

  // constant tables -> shared memory, once per persistent CTA
  for (int i = tid; i < kNfft; i += kWarps * 32) s_win[i] = p.window[i];
  for (int i = tid; i < 1024; i += kWarps * 32) s_tw2[i] = p.tw2[i];
  for (int i = tid; i < 513; i += kWarps * 32) s_twr[i] = p.twr[i];
  for (int i = tid; i < kMaxMelW; i += kWarps * 32) s_melw[i] = p.melw[i];
  for (int i = tid; i < kMelInfo; i += kWarps * 32) s_melinfo[i] = p.melinfo[i];
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  uint32_t phase0 = 0, phase1 = 0;
  int blk = blockIdx.x;
  if (blk < p.n_blocks && tid == 0) {
    BlockInfo b = resolve_block(p, blk);
    if (b.bulk) {
      mbar_expect_tx(&bars[0], (uint32_t)b.span * 4u);
      bulk_g2s(stage, b.clip.wav + b.s0, (uint32_t)b.span * 4u, &bars[0]);
    }
  }
  for (int it = 0; blk < p.n_blocks; blk += gridDim.x, ++it) {
    const int buf = it & 1;
    float* cur = stage + buf * kSpan;
    const BlockInfo b = resolve_block(p, blk);
    const int next = blk + gridDim.x;
    if (next < p.n_blocks && tid == 0) {          // prefetch the next block's samples behind this block's math
      BlockInfo nb = resolve_block(p, next);
      if (nb.bulk) {
        fence_proxy_async();
        mbar_expect_tx(&bars[buf ^ 1], (uint32_t)nb.span * 4u);
        bulk_g2s(stage + (buf ^ 1) * kSpan, nb.clip.wav + nb.s0, (uint32_t)nb.span * 4u, &bars[buf ^ 1]);
      }
    }
    if (b.bulk) {
      if (buf == 0) { mbar_wait(&bars[0], phase0); phase0 ^= 1; }
      else          { mbar_wait(&bars[1], phase1); phase1 ^= 1; }
    } else {                                       // clip edges / unaligned clips: guarded loads, zero "center" padding
      for (int i = tid; i < b.span; i += kWarps * 32) {
        long long s = b.s0 + i;
        cur[i] = (s >= 0 && s < b.clip.n_samples) ? __ldg(b.clip.wav + s) : 0.f;
      }
      __syncthreads();
    }
    for (int f = warp; f < b.nf; f += kWarps) {
      lm_rows(lane, cur + f * kHop, s_win, s_tw2, s_T);
      __syncwarp();
      float2 u[32];
      lm_cols_load(lane, s_T, u);
      __syncwarp();
      lm_cols_store(lane, u, s_T);
      __syncwarp();
      {
        float plo[16], phi[16], p0, p1024;
        lm_power_regs(lane, s_T, s_twr, plo, phi, p0, p1024);
        __syncwarp();                              // every lane has read its Z entries: P may overwrite them
        lm_power_store(lane, plo, phi, p0, p1024, s_P);
      }
      __syncwarp();
      lm_mel(lane, s_P, s_melw, s_melinfo, p.log_offset, b.clip.out + (b.t0 + f) * kNmels);
      __syncwarp();
    }
    __syncthreads();                               // stage[buf] is refilled two iterations from now
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
struct LogmelPlan {
  float* d_window = nullptr;
  float2* d_tw2 = nullptr;
  float2* d_twr = nullptr;
  float* d_melw = nullptr;
  uint32_t* d_melinfo = nullptr;
  float log_offset = 1e-8f;
  int sms = 148;
  // batch descriptors (device) + host staging for the host-buffer entry point
  LogmelClip* d_clips = nullptr;
  int* d_prefix = nullptr;
  int clip_cap = 0;
  float* d_wav = nullptr;
  float* d_out = nullptr;
  size_t wav_cap = 0, out_cap = 0;
};

static void default_window(std::vector<float>& w) {
  w.resize(kNfft);
  for (int n = 0; n < kNfft; ++n) w[n] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * (double)n / (double)kNfft));
}

// HTK-mel / slaney-normalised triangles for 16 kHz, 0..8000 Hz (torchaudio melscale_fbanks).  Computed in double;
// callers that need bit parity with torchaudio pass the table they built with torch ops instead.
static void default_fb(std::vector<float>& fb) {
  fb.assign((size_t)kNfreq * kNmels, 0.f);
  std::vector<double> f_pts(kNmels + 2);
  double m_max = 2595.0 * log10(1.0 + 8000.0 / 700.0);
  for (int i = 0; i < kNmels + 2; ++i) f_pts[i] = 700.0 * (pow(10.0, (m_max * i / (kNmels + 1)) / 2595.0) - 1.0);
  for (int k = 0; k < kNfreq; ++k) {
    double f = 8000.0 * k / (kNfreq - 1);
    for (int m = 0; m < kNmels; ++m) {
      double down = (f - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
      double up = (f_pts[m + 2] - f) / (f_pts[m + 2] - f_pts[m + 1]);
      double v = fmin(down, up);
      if (v > 0.0) fb[(size_t)k * kNmels + m] = (float)(v * 2.0 / (f_pts[m + 2] - f_pts[m]));
    }
  }
}

}  // namespace hft

using namespace hft;

extern "C" int hft_logmel_create(hft_logmel_plan** out, const float* window_host, const float* fb_host, float log_offset) {
  HFT_REQUIRE(out != nullptr, HFT_ERR_ARG, "hft_logmel_create: out is NULL");
  std::vector<float> win, fb;
  if (window_host) win.assign(window_host, window_host + kNfft); else default_window(win);
  if (fb_host) fb.assign(fb_host, fb_host + (size_t)kNfreq * kNmels); else default_fb(fb);
  // pack the filterbank into lane-major band weights (logmel_core.cuh)
  std::vector<float> melw(kMaxMelW, 0.f);
  std::vector<uint32_t> melinfo(kMelInfo);
  HFT_REQUIRE(lm_pack_filterbank(fb.data(), melw.data(), melinfo.data()) == 0, HFT_ERR_UNSUPPORTED,
              "hft_logmel_create: the fused kernel supports mel bands of <= 31 FFT bins and <= %d taps summed over the 8 bin groups", kMaxMelTaps);
  std::vector<float2> tw2(1024), twr(513);
  for (int k2 = 0; k2 < 32; ++k2)
    for (int n1 = 0; n1 < 32; ++n1) {
      double a = -2.0 * M_PI * (double)(n1 * k2) / 1024.0;
      tw2[k2 * 32 + n1] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int k = 0; k <= 512; ++k) {
    double a = -2.0 * M_PI * (double)k / 2048.0;
    twr[k] = make_float2((float)cos(a), (float)sin(a));
  }
  LogmelPlan* pl = new LogmelPlan();
  pl->log_offset = log_offset;
  pl->sms = num_sms();
  HFT_CHECK_CUDA(cudaMalloc(&pl->d_window, kNfft * 4));
  HFT_CHECK_CUDA(cudaMalloc(&pl->d_tw2, 1024 * 8));
  HFT_CHECK_CUDA(cudaMalloc(&pl->d_twr, 513 * 8));
  HFT_CHECK_CUDA(cudaMalloc(&pl->d_melw, kMaxMelW * 4));
  HFT_CHECK_CUDA(cudaMalloc(&pl->d_melinfo, kMelInfo * 4));
  HFT_CHECK_CUDA(cudaMemcpy(pl->d_window, win.data(), kNfft * 4, cudaMemcpyHostToDevice));
  HFT_CHECK_CUDA(cudaMemcpy(pl->d_tw2, tw2.data(), 1024 * 8, cudaMemcpyHostToDevice));
  HFT_CHECK_CUDA(cudaMemcpy(pl->d_twr, twr.data(), 513 * 8, cudaMemcpyHostToDevice));
  HFT_CHECK_CUDA(cudaMemcpy(pl->d_melw, melw.data(), kMaxMelW * 4, cudaMemcpyHostToDevice));
  HFT_CHECK_CUDA(cudaMemcpy(pl->d_melinfo, melinfo.data(), kMelInfo * 4, cudaMemcpyHostToDevice));
  HFT_CHECK_CUDA(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout::total));
  *out = reinterpret_cast<hft_logmel_plan*>(pl);
  return HFT_OK;
}

extern "C" int hft_logmel_destroy(hft_logmel_plan* plan) {
  if (!plan) return HFT_OK;
  LogmelPlan* pl = reinterpret_cast<LogmelPlan*>(plan);
  cudaFree(pl->d_window); cudaFree(pl->d_tw2); cudaFree(pl->d_twr); cudaFree(pl->d_melw); cudaFree(pl->d_melinfo);
  cudaFree(pl->d_clips); cudaFree(pl->d_prefix); cudaFree(pl->d_wav); cudaFree(pl->d_out);
  delete pl;
  return HFT_OK;
}

extern "C" int64_t hft_logmel_num_frames(int64_t n_samples) { return n_samples < 0 ? 0 : 1 + n_samples / kHop; }

static int launch_logmel(LogmelPlan* pl, LogmelParams& p, cudaStream_t stream) {
  p.window = pl->d_window; p.tw2 = pl->d_tw2; p.twr = pl->d_twr; p.melw = pl->d_melw; p.melinfo = pl->d_melinfo;
  p.log_offset = pl->log_offset;
  reset_launch_count();
  if (p.n_blocks == 0) return HFT_OK;
  int grid = p.n_blocks < pl->sms ? p.n_blocks : pl->sms;
  {
    LaunchScope ls(HFT_KCLASS_LOGMEL, stream);
    logmel_kernel<<<grid, kWarps * 32, SmemLayout::total, stream>>>(p);
  }
  HFT_CHECK_CUDA(cudaGetLastError());
  return HFT_OK;
}

extern "C" int hft_logmel_f32(hft_logmel_plan* plan, const float* wav_dev, int64_t n_samples, float* out_dev, int64_t n_frames,
                              void* stream) {
  HFT_REQUIRE(plan != nullptr, HFT_ERR_ARG, "hft_logmel_f32: plan is NULL");
  HFT_REQUIRE(n_samples >= 0 && n_frames == hft_logmel_num_frames(n_samples), HFT_ERR_ARG,
              "hft_logmel_f32: n_frames=%lld does not match 1 + n_samples/256 = %lld", (long long)n_frames,
              (long long)hft_logmel_num_frames(n_samples));
  HFT_REQUIRE(out_dev != nullptr && (wav_dev != nullptr || n_samples == 0), HFT_ERR_ARG, "hft_logmel_f32: NULL buffer");
  LogmelPlan* pl = reinterpret_cast<LogmelPlan*>(plan);
  long long nb = (n_frames + kFramesPerBlock - 1) / kFramesPerBlock;
  HFT_REQUIRE(nb < (1ll << 31), HFT_ERR_UNSUPPORTED, "hft_logmel_f32: clip too long");
  LogmelParams p{};
  p.n_clips = 1;
  p.n_blocks = (int)nb;
  p.single = LogmelClip{wav_dev, (long long)n_samples, out_dev, (long long)n_frames};
  return launch_logmel(pl, p, (cudaStream_t)stream);
}

extern "C" int hft_logmel_batch_f32(hft_logmel_plan* plan, const float* wav_dev, const int64_t* clip_start_host, const int64_t* clip_len_host,
                                    int n_clips, float* out_dev, void* stream) {
  HFT_REQUIRE(plan != nullptr && clip_start_host != nullptr && clip_len_host != nullptr && n_clips >= 0, HFT_ERR_ARG,
              "hft_logmel_batch_f32: bad argument");
  if (n_clips == 0) return HFT_OK;
  LogmelPlan* pl = reinterpret_cast<LogmelPlan*>(plan);
  std::vector<LogmelClip> clips(n_clips);
  std::vector<int> prefix(n_clips + 1, 0);
  long long frame_off = 0, blocks = 0;
  for (int c = 0; c < n_clips; ++c) {
    long long n = clip_len_host[c];
    HFT_REQUIRE(n >= 0 && clip_start_host[c] >= 0, HFT_ERR_ARG, "hft_logmel_batch_f32: negative start/length at clip %d", c);
    long long T = hft_logmel_num_frames(n);
    clips[c] = LogmelClip{wav_dev + clip_start_host[c], n, out_dev + frame_off * kNmels, T};
    frame_off += T;
    blocks += (T + kFramesPerBlock - 1) / kFramesPerBlock;
    HFT_REQUIRE(blocks < (1ll << 31), HFT_ERR_UNSUPPORTED, "hft_logmel_batch_f32: batch too long");
    prefix[c + 1] = (int)blocks;
  }
  cudaStream_t st = (cudaStream_t)stream;
  LogmelParams p{};
  p.n_clips = n_clips;
  p.n_blocks = (int)blocks;
  if (n_clips == 1) {
    p.single = clips[0];
  } else {
    if (pl->clip_cap < n_clips) {
      cudaFree(pl->d_clips); cudaFree(pl->d_prefix);
      HFT_CHECK_CUDA(cudaMalloc(&pl->d_clips, sizeof(LogmelClip) * n_clips));
      HFT_CHECK_CUDA(cudaMalloc(&pl->d_prefix, sizeof(int) * (n_clips + 1)));
      pl->clip_cap = n_clips;
    }
    // small descriptor upload; pageable source => the copy is staged before the call returns
    HFT_CHECK_CUDA(cudaMemcpyAsync(pl->d_clips, clips.data(), sizeof(LogmelClip) * n_clips, cudaMemcpyHostToDevice, st));
    HFT_CHECK_CUDA(cudaMemcpyAsync(pl->d_prefix, prefix.data(), sizeof(int) * (n_clips + 1), cudaMemcpyHostToDevice, st));
    p.clips = pl->d_clips;
    p.blk_prefix = pl->d_prefix;
  }
  return launch_logmel(pl, p, st);
}

extern "C" int hft_logmel_host_f32(hft_logmel_plan* plan, const float* wav_host, int64_t n_samples, float* out_host, int64_t n_frames,
                                   void* stream) {
  HFT_REQUIRE(plan != nullptr, HFT_ERR_ARG, "hft_logmel_host_f32: plan is NULL");
  HFT_REQUIRE(n_samples >= 0 && n_frames == hft_logmel_num_frames(n_samples), HFT_ERR_ARG, "hft_logmel_host_f32: n_frames mismatch");
  LogmelPlan* pl = reinterpret_cast<LogmelPlan*>(plan);
  cudaStream_t st = (cudaStream_t)stream;
  size_t wav_bytes = (size_t)n_samples * 4, out_bytes = (size_t)n_frames * kNmels * 4;
  if (pl->wav_cap < wav_bytes) {
    cudaFree(pl->d_wav);
    pl->wav_cap = 0;
    HFT_CHECK_CUDA(cudaMalloc(&pl->d_wav, wav_bytes + 16));
    pl->wav_cap = wav_bytes;
  }
  if (pl->out_cap < out_bytes) {
    cudaFree(pl->d_out);
    pl->out_cap = 0;
    HFT_CHECK_CUDA(cudaMalloc(&pl->d_out, out_bytes));
    pl->out_cap = out_bytes;
  }
  if (wav_bytes) HFT_CHECK_CUDA(cudaMemcpyAsync(pl->d_wav, wav_host, wav_bytes, cudaMemcpyHostToDevice, st));
  int rc = hft_logmel_f32(plan, pl->d_wav, n_samples, pl->d_out, n_frames, stream);
  if (rc != HFT_OK) return rc;
  HFT_CHECK_CUDA(cudaMemcpyAsync(out_host, pl->d_out, out_bytes, cudaMemcpyDeviceToHost, st));
  HFT_CHECK_CUDA(cudaStreamSynchronize(st));
  return HFT_OK;
}

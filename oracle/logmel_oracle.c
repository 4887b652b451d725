/* CPU oracle (plain C) for the log-mel front end  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path never links or calls it.
 *
 * Restates AMT.wav2feature for a mono 16 kHz waveform (reference hftt_code/model/amt.py:59-61,
 * hftt_code/corpus/config.json:2-12).  The arithmetic is torchaudio's (third-party, not under
 * /root/reference; pinned against torchaudio 2.11.0): see oracle/logmel_oracle.py for the published
 * algorithm.  fp32 arithmetic throughout, FFT-structured (radix-2), so that
 * the rounding behaviour is the log N kind of the reference's pocketfft path.
 *
 * Pinned by tests/test_oracle_logmel.py against tests/golden/logmel.npz (outputs of the reference).
 *
 * Build: make -C oracle      (gcc -O2 -pthread -shared -fPIC)
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define N_FFT 2048
#define HOP 256
#define N_FREQ 1025
#define N_MELS 256
#define LOG_N 11

typedef struct {
    float tw_re[N_FFT / 2], tw_im[N_FFT / 2];
    uint16_t rev[N_FFT];
} fft_tab;

static void fft_tab_init(fft_tab *t) {
    for (int k = 0; k < N_FFT / 2; ++k) {
        double a = -2.0 * M_PI * (double)k / (double)N_FFT;
        t->tw_re[k] = (float)cos(a);
        t->tw_im[k] = (float)sin(a);
    }
    for (int i = 0; i < N_FFT; ++i) {
        unsigned r = 0;
        for (int b = 0; b < LOG_N; ++b)
            if (i & (1 << b)) r |= 1u << (LOG_N - 1 - b);
        t->rev[i] = (uint16_t)r;
    }
}

/* in-place radix-2 decimation-in-time complex FFT of length 2048, fp32 */
static void fft2048(const fft_tab *t, float *re, float *im) {
    for (int len = 2; len <= N_FFT; len <<= 1) {
        int half = len >> 1, step = N_FFT / len;
        for (int i = 0; i < N_FFT; i += len) {
            for (int j = 0; j < half; ++j) {
                float wr = t->tw_re[j * step], wi = t->tw_im[j * step];
                int a = i + j, b = a + half;
                float xr = re[b] * wr - im[b] * wi;
                float xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr;
                im[b] = im[a] - xi;
                re[a] += xr;
                im[a] += xi;
            }
        }
    }
}

/* number of frames torch.stft(center=True) produces: amt.py:59 -> T = 1 + N / hop */
long oracle_logmel_num_frames(long n_samples) { return 1 + n_samples / HOP; }

typedef struct {
    const float *wav, *window, *fb;
    long n, t0, t1;
    float log_offset;
    float *out;
    const fft_tab *tab;
    const int *lo, *hi;
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    float *re = (float *)malloc(sizeof(float) * N_FFT);
    float *im = (float *)malloc(sizeof(float) * N_FFT);
    float *pw = (float *)malloc(sizeof(float) * N_FREQ);
    for (long t = j->t0; t < j->t1; ++t) {
        long base = t * HOP - N_FFT / 2; /* center=True, constant (zero) padding */
        for (int i = 0; i < N_FFT; ++i) {
            long s = base + i;
            float v = (s >= 0 && s < j->n) ? j->wav[s] * j->window[i] : 0.0f;
            int r = j->tab->rev[i];
            re[r] = v;
            im[r] = 0.0f;
        }
        fft2048(j->tab, re, im);
        for (int k = 0; k < N_FREQ; ++k) pw[k] = re[k] * re[k] + im[k] * im[k];
        float *o = j->out + t * N_MELS;
        for (int m = 0; m < N_MELS; ++m) {
            float acc = 0.0f;
            for (int k = j->lo[m]; k <= j->hi[m]; ++k) acc += pw[k] * j->fb[(long)k * N_MELS + m];
            o[m] = logf(acc + j->log_offset);
        }
    }
    free(re);
    free(im);
    free(pw);
    return NULL;
}

/* wav [n] fp32 mono 16 kHz; window [2048]; fb dense [1025][256] row-major (torchaudio layout);
 * out [T][256] fp32 = log(mel + log_offset)   (amt.py:61).  n_threads <= 64.  Returns 0. */
int oracle_logmel_f32(const float *wav, long n, const float *window, const float *fb, float log_offset,
                      float *out, int n_threads) {
    long T = oracle_logmel_num_frames(n);
    fft_tab *tab = (fft_tab *)malloc(sizeof(fft_tab));
    fft_tab_init(tab);
    /* band form of fb: first/last non-zero row per mel column */
    int lo[N_MELS], hi[N_MELS];
    for (int m = 0; m < N_MELS; ++m) {
        lo[m] = N_FREQ;
        hi[m] = -1;
        for (int k = 0; k < N_FREQ; ++k)
            if (fb[(long)k * N_MELS + m] != 0.0f) {
                if (k < lo[m]) lo[m] = k;
                hi[m] = k;
            }
    }
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    job_t jobs[64];
    pthread_t th[64];
    long per = (T + n_threads - 1) / n_threads;
    for (int i = 0; i < n_threads; ++i) {
        long t0 = i * per, t1 = t0 + per;
        if (t0 > T) t0 = T;
        if (t1 > T) t1 = T;
        jobs[i] = (job_t){wav, window, fb, n, t0, t1, log_offset, out, tab, lo, hi};
        if (i > 0) pthread_create(&th[i], NULL, worker, &jobs[i]);
    }
    worker(&jobs[0]);
    for (int i = 1; i < n_threads; ++i) pthread_join(th[i], NULL);
    free(tab);
    return 0;
}

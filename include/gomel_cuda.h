/*
 * gomel_cuda.h -- C ABI of libgomelcuda.so, the B200 (sm_100a) implementation of the gomel
 * spectrogram hot path.  Plain pointers and sizes only: this is what the reference's Go packages
 * bind through cgo (INTEGRATION.md shows the stub) and what gomel_b200/*.py bind through ctypes.
 *
 * The reference (neurlang/gomel) has no FFI of its own; each entry point below names the Go
 * function whose arithmetic it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every function returns 0 on success or a negative gomel_status; gomel_last_error(ctx)
 *     returns a message for the last failure on that context.
 *   - the caller allocates every output; the library never retains caller memory after return.
 *   - a context owns one device, one stream and grow-only scratch; calls on one context are
 *     serialised by an internal mutex and may come from any OS thread (cgo hops threads).
 *   - this build supports Resolut (n_fft) = 4096 and Window (hop) = 1280 (the configuration of
 *     cmd/tomel, cmd/towav, cmd/tophase, cmd/fromphase and of NewPhase) on every entry point,
 *     and Resolut 2048 / Window 256 (the mel.NewMel defaults, mel/mel.go:37-38) on the mel entry
 *     points (gomel_to_mel*, gomel_from_mel* except GOMEL_FLAG_F64_REF, gomel_set_mel_tables);
 *     anything else returns GOMEL_E_UNSUPPORTED.  There is NO CPU fallback.
 */
#ifndef GOMEL_CUDA_H
#define GOMEL_CUDA_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct gomel_ctx gomel_ctx;

typedef enum {
    GOMEL_OK = 0,
    GOMEL_E_ARG = -1,          /* bad argument, incl. len % n_mels != 0 where Go panics (mel/impl.go:366-372) */
    GOMEL_E_CUDA = -2,         /* CUDA runtime error */
    GOMEL_E_NOMEM = -3,        /* device or pinned allocation failed */
    GOMEL_E_UNSUPPORTED = -4,  /* Resolut / Window outside this build (see above) */
    GOMEL_E_STATE = -5         /* call order, e.g. mel tables not set */
} gomel_status;

/* The fields of mel.Mel (mel/mel.go:10-27) and phase.Phase (phase/phase.go:8-18) that reach
 * arithmetic.  YReverse / SampleRate / IHS / HDR only affect the PNG/WAV codecs on the host. */
typedef struct {
    int n_fft;           /* Resolut */
    int hop;             /* Window  */
    int n_mels;          /* NumMels */
    int n_freqs;         /* NumFreqs */
    int gl_iters;        /* GriffinLimIterations */
    double tune_mul;     /* TuneMul */
    double tune_add;     /* TuneAdd */
    double volume_boost; /* VolumeBoost (phase: multiplicative, applied iff != 0) */
    int flags;           /* GOMEL_FLAG_* */
    double mel_fmin;     /* MelFmin, MelFmax: with (n_fft, n_mels) the KEY of the filterbank tables registered by */
    double mel_fmax;     /* gomel_set_mel_tables -- never evaluated by the library.  Both 0: the most recently
                            registered tables for this (n_fft, n_mels). */
} gomel_config;

/* Precision of the Griffin-Lim loop (mel.ISTFT, mel/mel.go:76-139).  The loop is ill-conditioned while it has not
 * settled: a float32 rounding error made in iteration 0-1 ends ~300x larger after 32 iterations, and for about one
 * start signal in a hundred a trajectory passes near a singular point (a bin with a large target magnitude whose
 * analysis value is almost zero) somewhere in the first ~16 iterations, where one float32 rounding flips the
 * outcome by 1e-4 .. 1e-3.  Measured on 1,056 (clip, start signal) pairs at 32 iterations
 * (profiles/r02_gl_parity_sweep.md): an all-float32 loop misses the 1e-4 tolerance on 3 % of the pairs, 4 float64
 * lead iterations on 0.9 %, 12 on 0.1 %, 16 on none (worst 1.3e-5); on flat spectra float32 errors also grow in
 * late bursts, so long float32 tails are unsafe at 100 iterations.  (Later sweeps of 42,240 more pairs under the
 * default policy, profiles/r02_gl_guard.md: one pair at 2.7e-4 from a near-singular bin in the float32 tail -- now
 * re-run in float64 by the guard, gomel_set_gl_guard below -- and one at 1.3e-4 from a transient instability of the
 * iteration that only a shorter float32 tail removes: gomel_set_f32_tail(ctx, 8) holds all 10,560 pairs of that
 * population within 1.5e-5 at 15 % less throughput.)  Default policy: the first `lead` iterations run
 * in float64 end to end on the fused kernel of gl_f64.cuh, the rest in float32, with
 * lead = max(16, GriffinLimIterations - 16) -- at least sixteen float64 iterations first, at most sixteen float32
 * ones last (runs of <= 16 iterations are float64 throughout).  See gomel_set_lead_f64 / gomel_set_f32_tail.
 *   GOMEL_FLAG_F64      every iteration in float64 on the fused kernel (all from_mel entry points; the host-buffer
 *                       call then also reads the start signal and returns the waveform without a float32 step).
 *   GOMEL_FLAG_F64_REF  gomel_from_mel only: the round-1 strict path (one frame pair per CTA, spectra through HBM,
 *                       frames added in the reference's own order) -- a slow test instrument pinned < 1e-10 to
 *                       the oracle; the other two modes are checked against it. */
#define GOMEL_FLAG_F64 1
#define GOMEL_FLAG_F64_REF 2

/* ---- context ------------------------------------------------------------------------- */
int  gomel_ctx_create(int device, gomel_ctx **out);
void gomel_ctx_destroy(gomel_ctx *ctx);
const char *gomel_last_error(gomel_ctx *ctx);
const char *gomel_version(void);
/* number of kernels this context has launched so far (bench.py "gpu_launches") */
unsigned long long gomel_launch_count(gomel_ctx *ctx);
/* frames per tile for the tiled kernels; 0 = automatic (default) */
int  gomel_set_tile_frames(gomel_ctx *ctx, int tile_frames);
/* Griffin-Lim precision policy: lead = max(lead_iters, GriffinLimIterations - f32_tail) float64 iterations, then
 * float32.  Defaults 16 and 16 (env GOMEL_LEAD_F64, GOMEL_F32_TAIL); lead_iters = 0 with f32_tail < 0 (unlimited)
 * is the all-float32 loop of round 1.  Each returns the previous value (unlimited tail: INT_MAX). */
int  gomel_set_lead_f64(gomel_ctx *ctx, int lead_iters);
int  gomel_set_f32_tail(gomel_ctx *ctx, int f32_tail);
/* Singular-bin guard of the float32 tail.  A float32 iteration decides the phase of every bin from a value that
 * carries an absolute error of ~1e-7 of the frame's rms bin; when |X[k]| is that small while the target M[k] is not,
 * the float32 and float64 trajectories take different branches and end ~1e-4 apart -- about one (clip, start signal)
 * pair in 10,000 under the 16 + 16 policy (profiles/r02_gl_guard.md).  The float32 kernel therefore records, per
 * clip, the largest  leverage = M[k]/|X[k]| * rms_frame(M)/rms_clip(M)  it meets; clips above `threshold` have
 * their float32 iterations run again in float64 from the float64 signal of the hand-over, so that they end exactly
 * where GOMEL_FLAG_F64 ends.  Default 5e4 (env GOMEL_GL_GUARD): the one pair in 10,560 that missed 1e-4 had 1.3e6,
 * and the worst error a bin of leverage L can inject is about 2e-10 L; 0.5 % of clips are re-run.  0 disables.
 * The threshold is stated for 342-frame (10 s) clips and scaled by sqrt(frames / 342) inside: one bin's share of a
 * clip's norm falls with the square root of the frame count.  Not applied when no float64 lead iteration ran (lead_iters = 0) or on the
 * time-split sessions. */
int  gomel_set_gl_guard(gomel_ctx *ctx, float threshold, float *previous);
/* the guard's record of the last Griffin-Lim call on this context -- for the chunked *_batch_host calls, of their last
 * chunk -- (blocks until it has finished): clips seen (0: the guard did not run), clips re-run in float64, the largest
 * leverage, and the first `cap` clips' leverage */
int  gomel_last_gl_guard(gomel_ctx *ctx, int *n_clips, int *n_rerun, float *max_leverage, float *leverage, int cap);

/* ---- sizing: pad (mel/impl.go:429-455) + gossp NumFrames + ISTFT length (mel/mel.go:79) --- */
int  gomel_frames(const gomel_config *cfg, long n_samples, long *n_padded, long *n_frames, long *ola_len);
long gomel_ola_len(const gomel_config *cfg, long n_frames);

/* ---- mel filterbank tables ---------------------------------------------------------------
 * The per-band values domel (mel/impl.go:313-323) and undomel (:350-360) derive: int(inlo),
 * int(inhi), modlo.  Computed by the CALLER's math library (Go's math.Exp/Log in the cgo
 * binding) because two edges are 1-ulp fragile (SURVEY.md hard part 4); the library never
 * evaluates exp/log for table construction.  fwd_*: n_mels entries; inv_*: n_fft/2 entries.
 * A context keeps up to 64 table sets keyed by (n_fft, n_mels, mel_fmin, mel_fmax) of `cfg`; the mel
 * entry points look the set up by the same key of THEIR cfg, so goroutines / threads with different
 * Mel configurations can share one context (registering a key again replaces its tables). */
int  gomel_set_mel_tables(gomel_ctx *ctx, const gomel_config *cfg,
                          const int *fwd_lo, const int *fwd_hi, const double *fwd_mod,
                          const int *inv_lo, const int *inv_hi, const double *inv_mod);

/* ---- host-buffer API (what the Go / Python wrappers call; float64 like the reference) ---- */
/* mel.ToMel (mel/mel.go:46-74): wav[n] -> mel_out[frames*n_mels*2] */
int  gomel_to_mel(gomel_ctx *ctx, const gomel_config *cfg, const double *wav, long n, double *mel_out);
/* mel.FromMel (mel/mel.go:142-152): mel[n_frames*n_mels*2] -> wav_out[ola_len].
 * init_signal (ola_len doubles) replaces rand.Float64() of mel/mel.go:80-83; if NULL the
 * library fills U[0,1) on the device from `seed`.  The input is NOT modified: the wrapper
 * applies the reference's in-place exp side effect (mel/impl.go:421-427) itself. */
int  gomel_from_mel(gomel_ctx *ctx, const gomel_config *cfg, const double *mel, long n_frames,
                    const double *init_signal, unsigned long long seed, double *wav_out);
/* phase.ToPhase (phase/phase.go:41-70): wav[n] -> out[frames*n_freqs*2] = (Im X[j+1], Re X[j+1]) */
int  gomel_to_phase(gomel_ctx *ctx, const gomel_config *cfg, const double *wav, long n, double *out);
/* phase.FromPhase (phase/phase.go:136-153): spec[n_frames*n_freqs*2] -> wav_out[ola_len] */
int  gomel_from_phase(gomel_ctx *ctx, const gomel_config *cfg, const double *spec, long n_frames, double *wav_out);
/* Mel.Image / Phase.Image (mel/mel.go:171-173 -> mel/impl.go:16-44; phase/phase.go:190-192 ->
 * phase/impl.go:15-43): buf[n_entries*2] -> out[n_entries]; minmax_out (optional) = max0,max1,min0,min1 */
int  gomel_image(gomel_ctx *ctx, const double *buf, long n_entries, int mels, unsigned short *out, double *minmax_out);

/* ---- PNG pixel arithmetic (float64 on the device; the PNG/zlib container stays on the host) ----
 * gomel_quantise: the quantisation loops of mel dumpimage (mel/impl.go:138-181) and phase
 * dumpimage (phase/impl.go:170-266).  buf[n_entries*2] is left untouched; `ihs_passes` asinh
 * passes (phase/impl.go:171-177) are applied to a device copy first.
 *   flags: GOMEL_Q_SINGLE_MINMAX  one min/max over both channels (mel)  else per channel (phase)
 *          GOMEL_Q_HDR            maxVal 65535 (uint16 wrap) else 255 (uint8 wrap)
 *          GOMEL_Q_BLUE_WRAP      B = uintN(int(maxVal * (-val0)))  (phase/impl.go:229,256) else 0
 * rgb_out[n_entries*3] in buffer order (x*mels + y); minmax_out = max0,max1,min0,min1 (after asinh),
 * which the caller packs into the float16 metadata bytes. */
#define GOMEL_Q_SINGLE_MINMAX 1
#define GOMEL_Q_HDR 2
#define GOMEL_Q_BLUE_WRAP 4
int  gomel_quantise(gomel_ctx *ctx, const double *buf, long n_entries, int mels, int flags, int ihs_passes,
                    unsigned short *rgb_out, double *minmax_out);
/* gomel_dequantise: loadpng (mel/impl.go:92-112, phase/impl.go:98-147): rg[n_entries*2] pixel
 * values (0..255 or 0..65535) -> out[n_entries*2] = px/maxVal*(max-min)+min, then `ihs_passes` sinh. */
int  gomel_dequantise(gomel_ctx *ctx, const unsigned short *rg, long n_entries, int hdr, double max0,
                      double max1, double min0, double min1, int ihs_passes, double *out);

/* ---- device-resident API (float32 buffers in HBM; no host synchronisation unless stated) ---- */
int  gomel_dev_malloc(gomel_ctx *ctx, size_t bytes, void **out);
int  gomel_dev_free(gomel_ctx *ctx, void *p);
int  gomel_host_malloc(gomel_ctx *ctx, size_t bytes, void **out);     /* pinned */
int  gomel_host_free(gomel_ctx *ctx, void *p);
int  gomel_copy_h2d(gomel_ctx *ctx, void *dst, const void *src, size_t bytes);   /* async on ctx stream */
int  gomel_copy_d2h(gomel_ctx *ctx, void *dst, const void *src, size_t bytes);   /* async on ctx stream */
int  gomel_sync(gomel_ctx *ctx);
/* CUDA-event timer on the context's stream (the stream every kernel of the context runs on) */
int  gomel_timer_start(gomel_ctx *ctx);
int  gomel_timer_stop(gomel_ctx *ctx, float *ms);                     /* records, synchronises */
/* device time of the dominant kernel of the LAST transform issued on this context (all
 * Griffin-Lim iteration launches of a FromMel, or the STFT launch of a ToMel/ToPhase), measured
 * by CUDA events recorded around those launches on the context's stream; *launches = how many. */
int  gomel_last_hot_kernel_ms(gomel_ctx *ctx, float *ms, int *launches);
/* same for the float64 lead iterations of the last Griffin-Lim run (*launches = 0, *ms = 0 if there were none) */
int  gomel_last_lead_kernel_ms(gomel_ctx *ctx, float *ms, int *launches);

/* Batched STFT + mel (K1+K2).  d_sig: [n_clips][sig_stride] float32, each clip zero padded to
 * n_padded = sig_len samples; d_mel: [n_clips][n_frames][n_mels][2] float32 (natural log). */
int  gomel_to_mel_dev(gomel_ctx *ctx, const gomel_config *cfg, const float *d_sig, int n_clips,
                      long sig_stride, long sig_len, long n_frames, float *d_mel);
/* Batched STFT + phase representation (K1+K4). d_out: [n_clips][n_frames][n_freqs][2] */
int  gomel_to_phase_dev(gomel_ctx *ctx, const gomel_config *cfg, const float *d_sig, int n_clips,
                        long sig_stride, long sig_len, long n_frames, float *d_out);
/* Batched half spectra (tests): d_spec [n_clips][n_frames][2049][2] (Re, Im) */
int  gomel_stft_dev(gomel_ctx *ctx, const gomel_config *cfg, const float *d_sig, int n_clips,
                    long sig_stride, long sig_len, long n_frames, float *d_spec);
/* Batched FromMel (K3 + iters x K5).  d_mel as above; d_init/d_out: [n_clips][sig_stride] with
 * sig_stride >= ola_len; d_init may be NULL (device U[0,1) from seed).  d_init is not modified. */
int  gomel_from_mel_dev(gomel_ctx *ctx, const gomel_config *cfg, const float *d_mel, int n_clips,
                        long n_frames, const float *d_init, unsigned long long seed,
                        long sig_stride, float *d_out);
/* Batched FromPhase (K4+K6). d_spec: [n_clips][n_frames][n_freqs][2]; d_out: [n_clips][sig_stride] */
int  gomel_from_phase_dev(gomel_ctx *ctx, const gomel_config *cfg, const float *d_spec, int n_clips,
                          long n_frames, long sig_stride, float *d_out);

/* ---- one long clip split by time across ranks (config 5) -------------------------------------
 * Rank `rank` of `world` owns a contiguous, tile-aligned range of the clip's frames and keeps its
 * part of the signal and of the target magnitudes resident for all Griffin-Lim iterations.  Per
 * iteration and per rank boundary exactly two partial sums of Resolut-Window = 2816 floats cross:
 * the earlier rank's TAIL partial (-> next rank) and the later rank's HEAD partial (-> previous
 * rank); each side adds local + received (a+b == b+a, so both hold identical samples).
 * Two ways to move them: the library's own NCCL calls (gomel_ts_nccl_init / gomel_ts_run_nccl below, libnccl
 * dlopen()ed), or the caller's collective library -- the session exposes the four device pointers and a
 * communication stream and orders that stream against its kernels with events (gomel_ts_comm_begin / _end;
 * torch.distributed NCCL in gomel_b200/timesplit.py).
 *   part: 0 = all tiles in one launch; 1 = only the tiles that touch a rank boundary (own stream);
 *         2 = the interior tiles -- lets the exchange of iteration i overlap its interior work. */
typedef struct gomel_ts gomel_ts;
int  gomel_ts_create(gomel_ctx *ctx, const gomel_config *cfg, long n_frames_total, int rank, int world,
                     int tile_frames, gomel_ts **out);
/* as gomel_ts_create, plus `edge_frames`: length of the tiles next to a rank boundary (even, >= 4; 0 = uniform
 * tiles).  Short boundary tiles keep the exchange's critical path short while interior tiles stay long. */
int  gomel_ts_create2(gomel_ctx *ctx, const gomel_config *cfg, long n_frames_total, int rank, int world,
                      int tile_frames, int edge_frames, gomel_ts **out);
void gomel_ts_destroy(gomel_ts *ts);
/* frames [frame_begin, frame_begin+n_frames_local); local buffers hold global samples
 * [sample_begin, sample_begin+n_samples_local), n_samples_local = n_frames_local*Window + 2816 */
int  gomel_ts_range(gomel_ts *ts, long *frame_begin, long *n_frames_local, long *sample_begin, long *n_samples_local);
/* d_mel_local: [n_frames_local][n_mels][2] float32; d_init_local: n_samples_local floats or NULL
 * (device U[0,1) keyed by the GLOBAL sample index, so the result does not depend on the split) */
int  gomel_ts_load(gomel_ts *ts, const float *d_mel_local, const float *d_init_local, unsigned long long seed);
int  gomel_ts_iterate(gomel_ts *ts, int iter, int part);
int  gomel_ts_halo_ptrs(gomel_ts *ts, int iter, float **send_tail, float **send_head, float **recv_tail, float **recv_head);
/* The Griffin-Lim precision policy applies per iteration index: iterations [0, gomel_ts_lead_iters) run in float64
 * (lead from the session's cfg.gl_iters; GOMEL_FLAG_F64: all) and exchange 2816 DOUBLES per partial -- the four
 * pointers of gomel_ts_halo_ptrs then address float64 buffers; gomel_ts_halo_elem_bytes(iter) says which (8 / 4). */
int  gomel_ts_halo_elem_bytes(gomel_ts *ts, int iter);
int  gomel_ts_lead_iters(gomel_ts *ts);
void *gomel_ts_comm_stream(gomel_ts *ts);                 /* cudaStream_t */
int  gomel_ts_comm_begin(gomel_ts *ts, int iter);          /* comm stream waits for iteration iter's boundary tiles */
int  gomel_ts_comm_end(gomel_ts *ts, int iter);            /* marks the exchange of iteration iter done */
/* folds the last partials in and copies the local signal (n_samples_local floats) to d_out_local */
int  gomel_ts_finish(gomel_ts *ts, int iters, float *d_out_local);
int  gomel_ts_sync(gomel_ts *ts);                          /* host waits for all three streams of the session */
int  gomel_copy_d2d(gomel_ctx *ctx, void *dst, const void *src, size_t bytes, void *stream /* NULL = ctx stream */);
/* Optional in-library exchange: NCCL is dlopen()ed (libnccl.so.2 or $GOMEL_NCCL_LIB), the communicator is built
 * from a 128-byte ncclUniqueId the caller distributes (rank 0 creates it, every rank passes the same bytes), and
 * gomel_ts_run_nccl enqueues `n_iters` complete iterations -- boundary tiles, ncclSend/ncclRecv of the two
 * 2816-float partials per neighbour on the communication stream, interior tiles -- without returning to the
 * caller in between (no per-iteration host work beyond the launches). */
int  gomel_nccl_unique_id(gomel_ctx *ctx, char id_out[128]);
int  gomel_ts_nccl_init(gomel_ts *ts, const char id[128]);
int  gomel_ts_run_nccl(gomel_ts *ts, int first_iter, int n_iters, int overlap);

/* phase.ISTFT (phase/phase.go:93-133) of one long clip split by time over the same sessions (cfg.n_freqs,
 * cfg.volume_boost; uniform tiles): every rank inverts its own frames (gomel_ts_phase_istft; d_spec_local:
 * [n_frames_local][n_freqs][2] float32), ONE partial of 2816 floats per rank boundary travels from the earlier
 * rank's tail (send_tail) into the later rank's head region (recv_tail) -- by the caller on the communication stream
 * between gomel_ts_comm_begin / gomel_ts_comm_end, or by gomel_ts_phase_run_nccl -- and gomel_ts_phase_finish adds
 * it and applies the window-sum gain, a function of the GLOBAL sample index (max of the window sum in closed form
 * from host tables).  d_out_local: n_samples_local floats; samples [0, n_frames_local*Window) are this rank's. */
int  gomel_ts_phase_istft(gomel_ts *ts, const float *d_spec_local);
int  gomel_ts_phase_halo_ptrs(gomel_ts *ts, float **send_tail, float **recv_tail);
int  gomel_ts_phase_finish(gomel_ts *ts, float *d_out_local);
int  gomel_ts_phase_run_nccl(gomel_ts *ts, const float *d_spec_local, float *d_out_local);

/* ---- pipelined host batch (end-to-end: pinned host float32 in, float32 out, H2D/compute/D2H
 * overlapped chunk by chunk on three streams).  mel: [n_clips][n_frames*n_mels*2],
 * init: [n_clips][ola_len] or NULL, out: [n_clips][ola_len]. */
int  gomel_from_mel_batch_host(gomel_ctx *ctx, const gomel_config *cfg, const float *mel, int n_clips,
                               long n_frames, const float *init, unsigned long long seed, float *out,
                               int clips_per_chunk);
/* Same pipeline, output as the 16-bit PCM samples dumpwav writes (mel/impl.go:195-232: beep wav.Encode with
 * Precision 2 clamps to [-1,1] and converts int16(v * 32767)): out: int16 [n_clips][ola_len].  What cmd/towav
 * ends in; halves the device-to-host traffic of the float32 form. */
int  gomel_from_mel_batch_host_pcm16(gomel_ctx *ctx, const gomel_config *cfg, const float *mel, int n_clips,
                                     long n_frames, const float *init, unsigned long long seed, short *out,
                                     int clips_per_chunk);
/* wav: [n_clips][n_samples] -> mel_out: [n_clips][n_frames*n_mels*2] */
int  gomel_to_mel_batch_host(gomel_ctx *ctx, const gomel_config *cfg, const float *wav, int n_clips,
                             long n_samples, float *mel_out, int clips_per_chunk);

#ifdef __cplusplus
}
#endif
#endif

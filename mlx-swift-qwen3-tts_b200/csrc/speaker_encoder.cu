// ECAPA-TDNN speaker encoder (see speaker_encoder.h).
#include "speaker_encoder.h"

#include <algorithm>
#include <cmath>

namespace q3 {

namespace {

constexpr int kNfft = 1024, kHop = 256, kBins = kNfft / 2 + 1, kMels = 128;
constexpr int kKernels[5] = {5, 3, 3, 3, 1};    // SpeakerEncoderConfig defaults (:399-418): the reference builds no other configuration
constexpr int kDilations[5] = {1, 2, 3, 4, 1};
constexpr int kScale = 8;
constexpr float kEps = 1e-12f;
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_RELU_TANH = 2, ACT_SIGMOID = 3 };

// `reflectPadSignal` / `reflectPad1d` (:148-167, :213-232): index i of the padded sequence -> source index
__device__ __forceinline__ int reflect_src(int i, int pad, int n) {
  int s = i - pad;
  if (s < 0) s = -s;                       // pad .. 1
  else if (s >= n) s = 2 * (n - 1) - s;    // n-2, n-3, ...
  return s < 0 ? 0 : (s >= n ? n - 1 : s);
}

// |rfft(frame * window)| with a table-driven DFT: one thread per (frame, bin).  k * n mod 1024 indexes exact fp32 roundings of the
// double-precision twiddles, so the only error is the fp32 summation (1024 terms)
__global__ void __launch_bounds__(256) spk_stft_mag_kernel(const float* __restrict__ audio, int n, int frames, const float* __restrict__ cs,
                                                         const float* __restrict__ sn, const float* __restrict__ window, float* __restrict__ mag) {
  __shared__ float fr[kNfft];
  const int f = blockIdx.x;
  for (int j = threadIdx.x; j < kNfft; j += blockDim.x) fr[j] = audio[reflect_src(f * kHop + j, kNfft / 2, n)] * window[j];
  __syncthreads();
  for (int k = threadIdx.x; k < kBins; k += blockDim.x) {
    float re = 0.f, im = 0.f;
    for (int j = 0; j < kNfft; ++j) {
      const int idx = (k * j) & (kNfft - 1);
      re = fmaf(fr[j], cs[idx], re);
      im = fmaf(fr[j], sn[idx], im);
    }
    mag[(size_t)f * kBins + k] = sqrtf(re * re + im * im);
  }
}

// log(clip(mag . filterbank, 1e-5)) (:66-69): one thread per (frame, mel)
__global__ void spk_mel_kernel(const float* __restrict__ mag, const float* __restrict__ fb /*[513][128]*/, int frames, float* __restrict__ mel) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= frames * kMels) return;
  const int f = i / kMels, m = i - f * kMels;
  float acc = 0.f;
  for (int k = 0; k < kBins; ++k) acc = fmaf(mag[(size_t)f * kBins + k], fb[k * kMels + m], acc);
  mel[i] = logf(fmaxf(acc, 1e-5f));
}

// TimeDelayNetBlock (:234-257) on channels-last rows: y[t][yc0 + n] = act(b[n] + sum_tap sum_c in[reflect(t + tap*d)][c] * W[tap][n][c]),
// in[t][c] = x[t][xc0 + c] (+ add[t][ac0 + c]: Res2Net's `chunk + outputPart`, :293).  One warp per output value, lanes over c.
__global__ void __launch_bounds__(256) spk_tdnn_kernel(const float* __restrict__ x, int ldx, int xc0, const float* __restrict__ add, int lda, int ac0,
                                                     const float* __restrict__ w, const float* __restrict__ b, int T, int cin, int cout, int k, int dil,
                                                     float* __restrict__ y, int ldy, int yc0, int act) {
  const size_t wid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= (size_t)T * cout) return;
  const int t = (int)(wid / cout), n = (int)(wid - (size_t)t * cout);
  const int pad = (k - 1) * dil / 2;
  float acc = 0.f;
  for (int tap = 0; tap < k; ++tap) {
    const int ts = reflect_src(t + tap * dil, pad, T);
    const float* xr = x + (size_t)ts * ldx + xc0;
    const float* ar = add ? add + (size_t)ts * lda + ac0 : nullptr;
    const float* wr = w + ((size_t)tap * cout + n) * cin;
    for (int c = lane; c < cin; c += 32) acc = fmaf(ar ? xr[c] + ar[c] : xr[c], wr[c], acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    float v = acc + (b ? b[n] : 0.f);
    if (act == ACT_RELU) v = fmaxf(v, 0.f);
    else if (act == ACT_RELU_TANH) v = tanhf(fmaxf(v, 0.f));
    else if (act == ACT_SIGMOID) v = 1.0f / (1.0f + expf(-v));
    y[(size_t)t * ldy + yc0 + n] = v;
  }
}

__global__ void spk_copy_cols_kernel(const float* __restrict__ x, int ldx, int xc0, float* __restrict__ y, int ldy, int yc0, int T, int ncols) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * ncols) return;
  const int t = i / ncols, c = i - t * ncols;
  y[(size_t)t * ldy + yc0 + c] = x[(size_t)t * ldx + xc0 + c];
}

// per-channel mean over time and sqrt(var + eps) (population variance, :369-371); one warp per channel
__global__ void spk_time_stats_kernel(const float* __restrict__ x, int ld, int T, int C, float* __restrict__ mean, float* __restrict__ stdv) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  float s = 0.f;
  for (int t = lane; t < T; t += 32) s += x[(size_t)t * ld + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float m = s / (float)T;
  float v = 0.f;
  for (int t = lane; t < T; t += 32) { const float d = x[(size_t)t * ld + c] - m; v = fmaf(d, d, v); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) {
    mean[c] = m;
    if (stdv) stdv[c] = sqrtf(v / (float)T + kEps);
  }
}

// SqueezeExcitationRes2NetBlock tail (:318-321, :350-351): y[t][yc0 + c] = h[t][c] * se[c] + res[t][c]
__global__ void spk_scale_add_kernel(const float* __restrict__ h, const float* __restrict__ se, const float* __restrict__ res, int T, int C,
                                     float* __restrict__ y, int ldy, int yc0, float* __restrict__ y2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * C) return;
  const int t = i / C, c = i - t * C;
  const float v = fmaf(h[i], se[c], res[i]);
  y[(size_t)t * ldy + yc0 + c] = v;
  if (y2) y2[i] = v;  // dense copy: the next block's input / residual
}

// attention input of the pooling (:373-375): [x | mean | std] along channels
__global__ void spk_concat_stats_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ stdv, int T, int C,
                                        float* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * 3 * C) return;
  const int t = i / (3 * C), c = i - t * 3 * C;
  y[i] = c < C ? x[(size_t)t * C + c] : (c < 2 * C ? mean[c - C] : stdv[c - 2 * C]);
}

// softmax over time per channel, then weighted mean and weighted std (:385-394); one warp per channel.  pooled = [wmean (C) | wstd (C)]
__global__ void spk_asp_pool_kernel(const float* __restrict__ att, const float* __restrict__ x, int T, int C, float* __restrict__ pooled) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  float mx = -INFINITY;
  for (int t = lane; t < T; t += 32) mx = fmaxf(mx, att[(size_t)t * C + c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float den = 0.f, num = 0.f;
  for (int t = lane; t < T; t += 32) {
    const float e = expf(att[(size_t)t * C + c] - mx);
    den += e;
    num = fmaf(e, x[(size_t)t * C + c], num);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { den += __shfl_xor_sync(0xffffffffu, den, o); num += __shfl_xor_sync(0xffffffffu, num, o); }
  const float wm = num / den;
  float var = 0.f;
  for (int t = lane; t < T; t += 32) {
    const float e = expf(att[(size_t)t * C + c] - mx) / den, d = x[(size_t)t * C + c] - wm;
    var = fmaf(e, d * d, var);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  if (lane == 0) {
    pooled[c] = wm;
    pooled[C + c] = sqrtf(fmaxf(var, kEps));
  }
}

const char* kFirstKey = "speaker_encoder.blocks.0.conv.weight";

}  // namespace

bool SpeakerEncoderDev::present(const std::string& model_dir) {
  try {
    SafeTensors st(model_dir + "/model.safetensors");
    return st.tensors().count(kFirstKey) != 0;
  } catch (const Error&) {
    return false;
  }
}

SpeakerEncoderDev::Conv SpeakerEncoderDev::load_conv(const std::map<std::string, STensor>& t, const std::string& key) {
  auto it = t.find("speaker_encoder." + key + ".weight");
  auto ib = t.find("speaker_encoder." + key + ".bias");
  Q3_CHECK(it != t.end() && ib != t.end(), Q3TTS_ERR_BAD_WEIGHTS, "model.safetensors: missing speaker-encoder tensor '%s'", key.c_str());
  const STensor& s = it->second;
  Q3_CHECK(s.shape.size() == 3 && ib->second.numel() == s.shape[0], Q3TTS_ERR_BAD_WEIGHTS, "speaker-encoder conv '%s' has the wrong shape", key.c_str());
  Conv c;
  c.cout = (int)s.shape[0]; c.cin = (int)s.shape[1]; c.k = (int)s.shape[2];
  const std::vector<float> h = to_f32_host(s), hb = to_f32_host(ib->second);
  std::vector<float> r((size_t)c.k * c.cout * c.cin);  // disk [out][in][k] (PyTorch; `transposeConv` :544-548) -> [k][out][in]
  for (int o = 0; o < c.cout; ++o)
    for (int i = 0; i < c.cin; ++i)
      for (int kk = 0; kk < c.k; ++kk) r[((size_t)kk * c.cout + o) * c.cin + i] = h[((size_t)o * c.cin + i) * c.k + kk];
  float* dw = arena_.alloc_n<float>(r.size());
  float* db = arena_.alloc_n<float>(hb.size());
  Q3_CUDA(cudaMemcpy(dw, r.data(), r.size() * 4, cudaMemcpyHostToDevice));
  Q3_CUDA(cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
  c.w = dw; c.b = db;
  return c;
}

SpeakerEncoderDev::SpeakerEncoderDev(const std::string& model_dir, cudaStream_t stream, LaunchCounter* counter) : stream_(stream), counter_(counter) {
  SafeTensors st(model_dir + "/model.safetensors");
  const auto& t = st.tensors();
  block0_ = load_conv(t, "blocks.0.conv");
  Q3_CHECK(block0_.cin == kMels && block0_.k == kKernels[0], Q3TTS_ERR_BAD_WEIGHTS, "speaker encoder: blocks.0 must be a %d-tap conv over %d mel bins", kKernels[0], kMels);
  ch_ = block0_.cout;
  Q3_CHECK(ch_ % kScale == 0, Q3TTS_ERR_BAD_WEIGHTS, "speaker encoder: %d channels do not split into %d Res2Net chunks", ch_, kScale);
  for (int i = 0; i < 3; ++i) {
    const std::string p = "blocks." + std::to_string(i + 1);
    SEBlock& b = se_[i];
    b.tdnn1 = load_conv(t, p + ".tdnn1.conv");
    for (int j = 0; j < 7; ++j) b.res[j] = load_conv(t, p + ".res2net_block.blocks." + std::to_string(j) + ".conv");
    b.tdnn2 = load_conv(t, p + ".tdnn2.conv");
    b.se1 = load_conv(t, p + ".se_block.conv1");
    b.se2 = load_conv(t, p + ".se_block.conv2");
    Q3_CHECK(b.tdnn1.cin == ch_ && b.tdnn1.cout == ch_ && b.tdnn2.cin == ch_ && b.tdnn2.cout == ch_ && b.res[0].cin == ch_ / kScale &&
                 b.res[0].cout == ch_ / kScale && b.res[0].k == kKernels[i + 1] && b.se1.cin == ch_ && b.se2.cout == ch_ && b.se2.cin == b.se1.cout,
             Q3TTS_ERR_BAD_WEIGHTS, "speaker encoder: block %d has inconsistent shapes", i + 1);
  }
  mfa_ = load_conv(t, "mfa.conv");
  asp_tdnn_ = load_conv(t, "asp.tdnn.conv");
  asp_conv_ = load_conv(t, "asp.conv");
  fc_ = load_conv(t, "fc");
  mfa_ch_ = mfa_.cout;
  enc_dim_ = fc_.cout;
  Q3_CHECK(mfa_.cin == 3 * ch_ && mfa_.k == 1 && asp_tdnn_.cin == 3 * mfa_ch_ && asp_conv_.cin == asp_tdnn_.cout && asp_conv_.cout == mfa_ch_ &&
               fc_.cin == 2 * mfa_ch_ && fc_.k == 1,
           Q3TTS_ERR_BAD_WEIGHTS, "speaker encoder: pooling / output shapes are inconsistent");
  // front-end tables: DFT twiddles and the symmetric Hann window rounded once from double; Slaney mel filterbank in the reference's Float arithmetic
  std::vector<float> cs(kNfft), sn(kNfft), win(kNfft), fb((size_t)kBins * kMels, 0.f);
  for (int i = 0; i < kNfft; ++i) {
    cs[i] = (float)cos(2.0 * M_PI * i / kNfft);
    sn[i] = (float)-sin(2.0 * M_PI * i / kNfft);
    win[i] = 0.5f * (1.0f - cosf(2.0f * (float)M_PI * (float)i / (float)(kNfft - 1)));  // :181-183
  }
  {  // createMelFilterbankImpl (:75-146)
    const float f_sp = 200.0f / 3.0f, min_log_hz = 1000.0f, min_log_mel = min_log_hz / f_sp, log_step = (float)(log(6.4) / 27.0);
    auto hz_to_mel = [&](float hz) { return hz >= min_log_hz ? min_log_mel + (float)log((double)(hz / min_log_hz)) / log_step : hz / f_sp; };
    auto mel_to_hz = [&](float mel) { return mel >= min_log_mel ? min_log_hz * (float)exp((double)(log_step * (mel - min_log_mel))) : f_sp * mel; };
    const float m_min = hz_to_mel(0.0f), m_max = hz_to_mel(12000.0f);
    std::vector<float> f_pts(kMels + 2), f_diff(kMels + 1);
    for (int i = 0; i < kMels + 2; ++i) f_pts[i] = mel_to_hz(m_min + (float)i * (m_max - m_min) / (float)(kMels + 1));
    for (int i = 0; i < kMels + 1; ++i) f_diff[i] = f_pts[i + 1] - f_pts[i];
    for (int k = 0; k < kBins; ++k) {
      const float freq = (float)k * (float)(24000 / 2) / (float)(kBins - 1);
      for (int m = 0; m < kMels; ++m) {
        const float down = (freq - f_pts[m]) / f_diff[m], up = (f_pts[m + 2] - freq) / f_diff[m + 1];
        fb[(size_t)k * kMels + m] = std::max(0.0f, std::min(down, up)) * (2.0f / (f_pts[m + 2] - f_pts[m]));
      }
    }
  }
  auto up = [&](const std::vector<float>& h) {
    float* d = arena_.alloc_n<float>(h.size());
    Q3_CUDA(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    return (const float*)d;
  };
  d_cos_ = up(cs); d_sin_ = up(sn); d_window_ = up(win); d_fb_ = up(fb);
}

SpeakerEncoderDev::~SpeakerEncoderDev() {
  if (ws_) cudaFree(ws_);
}

void SpeakerEncoderDev::ensure_workspace(int64_t n_samples, int frames) {
  const size_t need = (size_t)n_samples + (size_t)frames * (kBins + kMels + 4 * (size_t)ch_ + 3 * (size_t)ch_ + 4 * (size_t)mfa_ch_ + asp_tdnn_.cout) + 8 * (size_t)mfa_ch_ +
                      4 * (size_t)ch_ + enc_dim_ + 1024;
  if (need <= ws_floats_) return;
  Q3_CUDA(cudaStreamSynchronize(stream_));
  if (ws_) cudaFree(ws_);
  ws_floats_ = need + need / 4;
  ws_bytes_ = ws_floats_ * 4;
  Q3_CUDA(cudaMalloc(&ws_, ws_bytes_));
}

void SpeakerEncoderDev::extract(const float* h_audio, int64_t n_samples, float* h_embedding, float* h_mels) {
  // reflect padding indexes pad .. 1: the STFT needs more than n_fft / 2 samples, the dilation-4 TimeDelayNet more than 4 frames -- below
  // that the reference indexes out of range (:213-232)
  Q3_CHECK(n_samples >= kNfft && n_samples < (1ll << 28), Q3TTS_ERR_INVALID_ARG,
           "speaker encoder: %lld samples (reflect padding needs at least %d)", (long long)n_samples, kNfft);
  const int T = frames_for(n_samples);
  ensure_workspace(n_samples, T);
  const LaunchCtx c = ctx();
  float* p = ws_;
  auto take = [&](size_t n) { float* r = p; p += (n + 3) / 4 * 4; return r; };
  float* audio = take((size_t)n_samples);
  float* mag = take((size_t)T * kBins);
  float* mel = take((size_t)T * kMels);
  float* x = take((size_t)T * ch_);      // current block input / residual
  float* h1 = take((size_t)T * ch_);
  float* h2 = take((size_t)T * ch_);
  float* h3 = take((size_t)T * ch_);
  float* cat = take((size_t)T * 3 * ch_);  // hiddenStatesList[1...] concatenated (:515)
  float* m = take((size_t)T * mfa_ch_);
  float* att_in = take((size_t)T * 3 * mfa_ch_);
  float* att_h = take((size_t)T * asp_tdnn_.cout);
  float* att = take((size_t)T * mfa_ch_);
  float* mean = take((size_t)mfa_ch_ + ch_);
  float* stdv = take((size_t)mfa_ch_);
  float* se_a = take((size_t)ch_);
  float* se_b = take((size_t)ch_);
  float* pooled = take((size_t)2 * mfa_ch_);
  float* emb = take((size_t)enc_dim_);
  Q3_CUDA(cudaMemcpyAsync(audio, h_audio, (size_t)n_samples * 4, cudaMemcpyHostToDevice, stream_));
  // ---- log-mel (:37-73)
  spk_stft_mag_kernel<<<T, 256, 0, stream_>>>(audio, (int)n_samples, T, d_cos_, d_sin_, d_window_, mag); c.tick();
  spk_mel_kernel<<<(T * kMels + 255) / 256, 256, 0, stream_>>>(mag, d_fb_, T, mel); c.tick();
  auto tdnn = [&](const Conv& w, const float* in, int ld_in, int c0_in, const float* add, int ld_add, int c0_add, int rows, int dil, float* out, int ld_out, int c0_out,
                  int act) {
    const size_t warps = (size_t)rows * w.cout;
    spk_tdnn_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, stream_>>>(in, ld_in, c0_in, add, ld_add, c0_add, w.w, w.b, rows, w.cin, w.cout, w.k, dil, out, ld_out,
                                                                             c0_out, act);
    c.tick();
  };
  // ---- blocks (:496-513)
  tdnn(block0_, mel, kMels, 0, nullptr, 0, 0, T, kDilations[0], x, ch_, 0, ACT_RELU);
  const int chunk = ch_ / kScale;
  for (int i = 0; i < 3; ++i) {
    const SEBlock& b = se_[i];
    const int dil = kDilations[i + 1];
    tdnn(b.tdnn1, x, ch_, 0, nullptr, 0, 0, T, 1, h1, ch_, 0, ACT_RELU);
    // Res2NetBlock (:282-301): chunk 0 passes through, chunk 1 = tdnn0(x1), chunk i = tdnn_{i-1}(x_i + y_{i-1})
    spk_copy_cols_kernel<<<(T * chunk + 255) / 256, 256, 0, stream_>>>(h1, ch_, 0, h2, ch_, 0, T, chunk); c.tick();
    for (int j = 1; j < kScale; ++j)
      tdnn(b.res[j - 1], h1, ch_, j * chunk, j >= 2 ? h2 : nullptr, ch_, (j - 1) * chunk, T, dil, h2, ch_, j * chunk, ACT_RELU);
    tdnn(b.tdnn2, h2, ch_, 0, nullptr, 0, 0, T, 1, h3, ch_, 0, ACT_RELU);
    // SqueezeExcitationBlock (:314-321)
    spk_time_stats_kernel<<<(ch_ * 32 + 255) / 256, 256, 0, stream_>>>(h3, ch_, T, ch_, mean, nullptr); c.tick();
    tdnn(b.se1, mean, ch_, 0, nullptr, 0, 0, 1, 1, se_a, b.se1.cout, 0, ACT_RELU);
    tdnn(b.se2, se_a, b.se1.cout, 0, nullptr, 0, 0, 1, 1, se_b, ch_, 0, ACT_SIGMOID);
    spk_scale_add_kernel<<<(T * ch_ + 255) / 256, 256, 0, stream_>>>(h3, se_b, x, T, ch_, cat, 3 * ch_, i * ch_, h1); c.tick();
    std::swap(x, h1);  // the block's output is the next block's input and residual
  }
  tdnn(mfa_, cat, 3 * ch_, 0, nullptr, 0, 0, T, kDilations[4], m, mfa_ch_, 0, ACT_RELU);
  // ---- AttentiveStatisticsPooling (:366-396)
  spk_time_stats_kernel<<<(mfa_ch_ * 32 + 255) / 256, 256, 0, stream_>>>(m, mfa_ch_, T, mfa_ch_, mean, stdv); c.tick();
  spk_concat_stats_kernel<<<(unsigned)(((size_t)T * 3 * mfa_ch_ + 255) / 256), 256, 0, stream_>>>(m, mean, stdv, T, mfa_ch_, att_in); c.tick();
  tdnn(asp_tdnn_, att_in, 3 * mfa_ch_, 0, nullptr, 0, 0, T, 1, att_h, asp_tdnn_.cout, 0, ACT_RELU_TANH);
  tdnn(asp_conv_, att_h, asp_tdnn_.cout, 0, nullptr, 0, 0, T, 1, att, mfa_ch_, 0, ACT_NONE);
  spk_asp_pool_kernel<<<(mfa_ch_ * 32 + 255) / 256, 256, 0, stream_>>>(att, m, T, mfa_ch_, pooled); c.tick();
  tdnn(fc_, pooled, 2 * mfa_ch_, 0, nullptr, 0, 0, 1, 1, emb, enc_dim_, 0, ACT_NONE);
  Q3_CUDA(cudaMemcpyAsync(h_embedding, emb, (size_t)enc_dim_ * 4, cudaMemcpyDeviceToHost, stream_));
  if (h_mels) Q3_CUDA(cudaMemcpyAsync(h_mels, mel, (size_t)T * kMels * 4, cudaMemcpyDeviceToHost, stream_));
  Q3_CUDA(cudaStreamSynchronize(stream_));
}

}  // namespace q3

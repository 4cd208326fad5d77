// ICL reference-audio encoder (see audio_encoder.h): weight loading in the on-disk (PyTorch) layouts + the encode graph of
// Qwen3TTSAudioEncoder.callAsFunction (Vocoder/Qwen3TTSAudioEncoder.swift:530-572).  fp32 end to end: the output is a table of
// nearest-neighbour indices, so every rounding that is not the reference's can flip a near-tie.
#include "audio_encoder.h"

#include <algorithm>
#include <cmath>

#include "codec_kernels.h"

namespace q3 {

namespace {
const STensor* find_t(const std::map<std::string, STensor>& t, const std::string& k) {
  auto it = t.find(k);
  return it == t.end() ? nullptr : &it->second;
}
const STensor& need_t(const std::map<std::string, STensor>& t, const std::string& k) {
  const STensor* s = find_t(t, k);
  if (!s) fail(Q3TTS_ERR_DECODER_LOAD_FAILED, "speech_tokenizer/model.safetensors: missing encoder tensor '%s'", k.c_str());
  return *s;
}
const char* kFirstKey = "encoder.encoder.layers.0.conv.weight";

Json load_tokenizer_config(const std::string& dir) {
  for (const char* name : {"config.json", "configuration.json", "speech_tokenizer_config.json"}) {
    const std::string p = dir + "/" + name;
    if (FILE* f = fopen(p.c_str(), "rb")) {
      fclose(f);
      return parse_json_file(p);
    }
  }
  return Json();
}
}  // namespace

bool AudioEncoder::present(const std::string& dir) {
  try {
    SafeTensors st(dir + "/model.safetensors");
    return st.tensors().count(kFirstKey) != 0;
  } catch (const Error&) {
    return false;
  }
}

const float* AudioEncoder::load_vec(const std::map<std::string, STensor>& t, const std::string& key, int n) {
  std::vector<float> h = to_f32_host(need_t(t, key));
  Q3_CHECK((int)h.size() == n, Q3TTS_ERR_DECODER_LOAD_FAILED, "encoder tensor '%s' has %zu elements, expected %d", key.c_str(), h.size(), n);
  float* d = arena_.alloc_n<float>(h.size());
  Q3_CUDA(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  return d;
}

// MimiConv1d (:24-85), disk [C_out, C_in, K].  stride 1: K causal taps, device [K][C_out][C_in].  stride r with K = 2r (every strided
// conv of this encoder): two taps over the [T / r, r * C_in] view of the input -- tap 0 = kernel indices [0, r) against the previous
// row, tap 1 = [r, 2r) against the current row (left padding K - r = r samples = exactly one row of the view).
ConvW AudioEncoder::load_conv(const std::map<std::string, STensor>& t, const std::string& key, int cout, int cin, int k, int stride) {
  const STensor& s = need_t(t, key + ".weight");
  Q3_CHECK(s.shape.size() == 3 && s.shape[0] == cout && s.shape[1] == cin && s.shape[2] == k, Q3TTS_ERR_DECODER_LOAD_FAILED,
           "encoder conv weight '%s' has the wrong shape", key.c_str());
  Q3_CHECK(stride == 1 || k == 2 * stride, Q3TTS_ERR_BAD_CONFIG, "encoder conv '%s': kernel %d / stride %d unsupported (strided convs need K = 2 * stride)",
           key.c_str(), k, stride);
  std::vector<float> h = to_f32_host(s);
  ConvW w;
  std::vector<float> r;
  if (stride == 1) {
    r.resize((size_t)k * cout * cin);
    for (int o = 0; o < cout; ++o)
      for (int i = 0; i < cin; ++i)
        for (int kk = 0; kk < k; ++kk) r[((size_t)kk * cout + o) * cin + i] = h[((size_t)o * cin + i) * k + kk];
    w.ntap = k; w.cin = cin;
  } else {
    const int cw = stride * cin;
    r.resize((size_t)2 * cout * cw);
    for (int tap = 0; tap < 2; ++tap)
      for (int o = 0; o < cout; ++o)
        for (int j = 0; j < stride; ++j)
          for (int i = 0; i < cin; ++i) r[((size_t)tap * cout + o) * cw + (size_t)j * cin + i] = h[((size_t)o * cin + i) * k + tap * stride + j];
    w.ntap = 2; w.cin = cw;
  }
  float* d = arena_.alloc_n<float>(r.size());
  Q3_CUDA(cudaMemcpy(d, r.data(), r.size() * 4, cudaMemcpyHostToDevice));
  w.w = d; w.dil = 1; w.n = cout;
  if (find_t(t, key + ".bias")) w.bias = load_vec(t, key + ".bias", cout);
  return w;
}

ConvW AudioEncoder::load_linear(const std::map<std::string, STensor>& t, const std::vector<std::string>& keys, const std::vector<int>& outs, int cin,
                                bool bias) {
  std::vector<float> all, ball;
  int n = 0;
  for (size_t i = 0; i < keys.size(); ++i) {
    const STensor& s = need_t(t, keys[i] + ".weight");
    Q3_CHECK(s.numel() == (int64_t)outs[i] * cin, Q3TTS_ERR_DECODER_LOAD_FAILED, "encoder linear '%s' has the wrong size", keys[i].c_str());
    std::vector<float> h = to_f32_host(s);
    all.insert(all.end(), h.begin(), h.end());
    if (bias) {
      std::vector<float> b = to_f32_host(need_t(t, keys[i] + ".bias"));
      Q3_CHECK((int)b.size() == outs[i], Q3TTS_ERR_DECODER_LOAD_FAILED, "bias of '%s' has the wrong size", keys[i].c_str());
      ball.insert(ball.end(), b.begin(), b.end());
    }
    n += outs[i];
  }
  float* d = arena_.alloc_n<float>(all.size());
  Q3_CUDA(cudaMemcpy(d, all.data(), all.size() * 4, cudaMemcpyHostToDevice));
  ConvW w;
  w.w = d; w.ntap = 1; w.dil = 1; w.cin = cin; w.n = n;
  if (bias) {
    float* db = arena_.alloc_n<float>(ball.size());
    Q3_CUDA(cudaMemcpy(db, ball.data(), ball.size() * 4, cudaMemcpyHostToDevice));
    w.bias = db;
  }
  return w;
}

AudioEncoder::AudioEncoder(const std::string& dir, cudaStream_t stream, LaunchCounter* counter) : stream_(stream), counter_(counter) {
  const Json root = load_tokenizer_config(dir);
  AudioEncoderConfig& c = cfg_;
  if (const Json* e = root.find("encoder_config")) {
    if (e->type == Json::Obj) {
      c.audio_channels = e->integer_or("audio_channels", c.audio_channels);
      c.codebook_size = e->integer_or("codebook_size", c.codebook_size);
      c.compress = e->integer_or("compress", c.compress);
      c.dilation_growth_rate = e->integer_or("dilation_growth_rate", c.dilation_growth_rate);
      c.hidden_size = e->integer_or("hidden_size", c.hidden_size);
      c.intermediate_size = e->integer_or("intermediate_size", c.intermediate_size);
      c.kernel_size = e->integer_or("kernel_size", c.kernel_size);
      c.last_kernel_size = e->integer_or("last_kernel_size", c.last_kernel_size);
      c.num_filters = e->integer_or("num_filters", c.num_filters);
      c.num_hidden_layers = e->integer_or("num_hidden_layers", c.num_hidden_layers);
      c.num_residual_layers = e->integer_or("num_residual_layers", c.num_residual_layers);
      c.num_quantizers = e->integer_or("num_quantizers", c.num_quantizers);
      c.num_semantic_quantizers = e->integer_or("num_semantic_quantizers", c.num_semantic_quantizers);
      c.upsampling_ratios = e->int_array_or("upsampling_ratios", c.upsampling_ratios);
      c.head_dim = e->integer_or("head_dim", c.head_dim);
      c.num_attention_heads = e->integer_or("num_attention_heads", c.num_attention_heads);
      c.num_key_value_heads = e->integer_or("num_key_value_heads", c.num_key_value_heads);
      c.norm_eps = (float)e->number_or("norm_eps", c.norm_eps);
      c.rope_theta = (float)e->number_or("rope_theta", c.rope_theta);
      c.vector_quantization_hidden_dimension = e->integer_or("vector_quantization_hidden_dimension", c.vector_quantization_hidden_dimension);
    }
  }
  c.valid_num_quantizers = root.integer_or("encoder_valid_num_quantizers", 16);
  Q3_CHECK(c.audio_channels == 1, Q3TTS_ERR_BAD_CONFIG, "audio encoder: mono input only (audio_channels %d)", c.audio_channels);
  Q3_CHECK(c.head_dim == 64, Q3TTS_ERR_BAD_CONFIG, "audio encoder attention kernels are specialised for head_dim 64 (got %d)", c.head_dim);
  Q3_CHECK(c.num_residual_layers == 1, Q3TTS_ERR_BAD_CONFIG, "audio encoder: num_residual_layers %d unsupported (1 only: dilation 1)", c.num_residual_layers);
  Q3_CHECK(c.num_attention_heads % c.num_key_value_heads == 0 && c.vector_quantization_hidden_dimension <= 1024, Q3TTS_ERR_BAD_CONFIG, "audio encoder: bad dims");
  SafeTensors st(dir + "/model.safetensors", Q3TTS_ERR_DECODER_LOAD_FAILED);
  const auto& t = st.tensors();
  const std::string E = "encoder.encoder.layers.";
  int li = 0;
  conv0_ = load_conv(t, E + std::to_string(li) + ".conv", c.num_filters, c.audio_channels, c.kernel_size, 1);
  ++li;
  int cur = c.num_filters;
  for (size_t i = 0; i < c.upsampling_ratios.size(); ++i) {
    const int ratio = c.upsampling_ratios[c.upsampling_ratios.size() - 1 - i];  // reversed (:135)
    Stage s;
    s.ratio = ratio; s.cin = cur; s.cout = c.num_filters << (i + 1);
    s.res1 = load_conv(t, E + std::to_string(li) + ".block.1.conv", cur / 2, cur, 3, 1);
    s.res2 = load_conv(t, E + std::to_string(li) + ".block.3.conv", cur, cur / 2, 1, 1);
    li += 2;  // the resnet block, then the ELU slot
    s.down = load_conv(t, E + std::to_string(li) + ".conv", s.cout, cur, 2 * ratio, ratio);
    ++li;
    cur = s.cout;
    stages_.push_back(s);
  }
  ++li;  // ELU slot
  conv_last_ = load_conv(t, E + std::to_string(li) + ".conv", c.hidden_size, cur, c.last_kernel_size, 1);
  const int H = c.hidden_size, qd = c.num_attention_heads * c.head_dim, kd = c.num_key_value_heads * c.head_dim;
  for (int n = 0; n < c.num_hidden_layers; ++n) {
    const std::string p = "encoder.encoder_transformer.layers." + std::to_string(n);
    TLayer l;
    l.qkv = load_linear(t, {p + ".self_attn.q_proj", p + ".self_attn.k_proj", p + ".self_attn.v_proj"}, {qd, kd, kd}, H, false);
    l.o = load_linear(t, {p + ".self_attn.o_proj"}, {H}, qd, false);
    l.fc1 = load_linear(t, {p + ".mlp.fc1"}, {c.intermediate_size}, H, true);
    l.fc2 = load_linear(t, {p + ".mlp.fc2"}, {H}, c.intermediate_size, true);
    l.ln1_w = load_vec(t, p + ".input_layernorm.weight", H); l.ln1_b = load_vec(t, p + ".input_layernorm.bias", H);
    l.ln2_w = load_vec(t, p + ".post_attention_layernorm.weight", H); l.ln2_b = load_vec(t, p + ".post_attention_layernorm.bias", H);
    l.attn_scale = load_vec(t, p + ".self_attn_layer_scale.scale", H);
    l.mlp_scale = load_vec(t, p + ".mlp_layer_scale.scale", H);
    tl_.push_back(l);
  }
  downsample_ = load_conv(t, "encoder.downsample.conv.conv", H, H, 2 * c.compress, c.compress);
  const int D = c.vector_quantization_hidden_dimension;
  n_sem_ = c.num_semantic_quantizers;
  n_out_ = std::min(c.valid_num_quantizers, c.num_quantizers);
  Q3_CHECK(n_out_ >= n_sem_ && n_out_ <= 64, Q3TTS_ERR_BAD_CONFIG, "audio encoder: bad quantizer counts");
  const std::string QS = "encoder.quantizer.semantic_residual_vector_quantizer", QA = "encoder.quantizer.acoustic_residual_vector_quantizer";
  {
    const STensor& ps = need_t(t, QS + ".input_proj.weight");
    const STensor& pa = need_t(t, QA + ".input_proj.weight");
    Q3_CHECK(ps.numel() == (int64_t)D * H && pa.numel() == (int64_t)D * H, Q3TTS_ERR_DECODER_LOAD_FAILED, "encoder quantizer input_proj has the wrong size");
  }
  proj_sem_ = load_linear(t, {QS + ".input_proj"}, {D}, H, false);  // Conv1d k = 1, no bias (:386)
  proj_ac_ = load_linear(t, {QA + ".input_proj"}, {D}, H, false);
  // codebook = embedding_sum / clip(cluster_usage, 1e-5) (:627-645); only the layers whose codes reach the output are resident
  for (int q = 0; q < n_out_; ++q) {
    const std::string p = (q < n_sem_ ? QS + ".layers." + std::to_string(q) : QA + ".layers." + std::to_string(q - n_sem_)) + "._codebook";
    std::vector<float> es = to_f32_host(need_t(t, p + ".embedding_sum")), cu = to_f32_host(need_t(t, p + ".cluster_usage"));
    Q3_CHECK((int)cu.size() == c.codebook_size && es.size() == (size_t)c.codebook_size * D, Q3TTS_ERR_DECODER_LOAD_FAILED, "encoder codebook '%s' has the wrong shape", p.c_str());
    std::vector<float> sq(c.codebook_size);
    for (int e = 0; e < c.codebook_size; ++e) {
      const float u = std::max(cu[e], 1e-5f);
      float acc = 0.f;
      for (int d = 0; d < D; ++d) {
        float& v = es[(size_t)e * D + d];
        v = v / u;
        acc += v * v;
      }
      sq[e] = acc;
    }
    float* db = arena_.alloc_n<float>(es.size());
    float* ds = arena_.alloc_n<float>(sq.size());
    Q3_CUDA(cudaMemcpy(db, es.data(), es.size() * 4, cudaMemcpyHostToDevice));
    Q3_CUDA(cudaMemcpy(ds, sq.data(), sq.size() * 4, cudaMemcpyHostToDevice));
    books_.push_back(db);
    books_sq_.push_back(ds);
  }
  d_books_ = (const float**)arena_.alloc(sizeof(float*) * n_out_);
  d_books_sq_ = (const float**)arena_.alloc(sizeof(float*) * n_out_);
  Q3_CUDA(cudaMemcpy(d_books_, books_.data(), sizeof(float*) * n_out_, cudaMemcpyHostToDevice));
  Q3_CUDA(cudaMemcpy(d_books_sq_, books_sq_.data(), sizeof(float*) * n_out_, cudaMemcpyHostToDevice));
  std::vector<float> f(32);
  for (int i = 0; i < 32; ++i) f[i] = 1.0f / powf(c.rope_theta, (float)(2 * i) / (float)c.head_dim);  // DecoderRotaryEmbedding (SpeechTokenizer.swift:286-287)
  float* df = arena_.alloc_n<float>(32);
  Q3_CUDA(cudaMemcpy(df, f.data(), 128, cudaMemcpyHostToDevice));
  d_inv_freq_ = df;
}

AudioEncoder::~AudioEncoder() {
  for (float*& p : ws_)
    if (p) { cudaFree(p); p = nullptr; }
  if (d_codes_) cudaFree(d_codes_);
}

int AudioEncoder::frames_for(int64_t n) const {
  int64_t T = n;
  for (size_t i = 0; i < cfg_.upsampling_ratios.size(); ++i) {
    const int r = cfg_.upsampling_ratios[i];
    T = (T + r - 1) / r;  // MimiConv1d with K = 2r: ceil(T / r) frames (extra right padding completes the last one, :55-62)
  }
  return (int)((T + cfg_.compress - 1) / cfg_.compress);
}

void AudioEncoder::ensure_workspace(int64_t samples) {
  if (samples <= ws_samples_) return;
  samples = (samples + 65535) / 65536 * 65536;
  const AudioEncoderConfig& c = cfg_;
  // widest activation: [L, num_filters] right after the first conv (channels double while time shrinks by >= 4); the transformer's
  // [T, max(3 * heads * 64, intermediate)] at T = L / 960 is far smaller
  int64_t stride_all = 1;
  for (int r : c.upsampling_ratios) stride_all *= r;
  const int64_t T = samples / stride_all + 2;
  const int64_t wide = std::max<int64_t>((int64_t)(c.num_attention_heads + 2 * c.num_key_value_heads) * c.head_dim, c.intermediate_size);
  const size_t floats = (size_t)std::max<int64_t>((samples + 64) * c.num_filters, T * wide + 1024);
  Q3_CUDA(cudaStreamSynchronize(stream_));
  for (float*& p : ws_) {
    if (p) cudaFree(p);
    p = nullptr;
    Q3_CUDA(cudaMalloc(&p, floats * sizeof(float)));
  }
  if (d_codes_) cudaFree(d_codes_);
  d_codes_ = nullptr;
  Q3_CUDA(cudaMalloc(&d_codes_, sizeof(int32_t) * (size_t)n_out_ * (size_t)T));
  ws_floats_ = floats;
  ws_bytes_ = 3 * floats * sizeof(float);
  ws_samples_ = samples;
}

int AudioEncoder::encode(const float* h_audio, int64_t n_samples, int32_t* h_codes, int capacity_frames, float* h_latent) {
  const AudioEncoderConfig& c = cfg_;
  Q3_CHECK(h_audio != nullptr && n_samples >= 1, Q3TTS_ERR_INVALID_ARG, "encode: no audio");
  Q3_CHECK(n_samples <= (int64_t)24000 * 180, Q3TTS_ERR_CAPACITY, "encode: reference audio longer than 180 s");
  const int frames = frames_for(n_samples);
  Q3_CHECK(frames <= capacity_frames, Q3TTS_ERR_CAPACITY, "encode: %d frames do not fit the caller's %d-frame buffer", frames, capacity_frames);
  ensure_workspace(n_samples);
  const LaunchCtx lc = ctx();
  float *A = ws_[0], *B = ws_[1], *C = ws_[2];
  Q3_CUDA(cudaMemcpyAsync(B, h_audio, sizeof(float) * (size_t)n_samples, cudaMemcpyHostToDevice, stream_));
  int64_t T = n_samples;
  // SEANet CNN (:120-190)
  launch_conv_gemm(lc, B, conv0_, A, nullptr, nullptr, (int)T, (int)T, CE_STORE);
  for (const Stage& s : stages_) {
    // MimiResnetBlock (:89-116): h + conv1(ELU(conv3(ELU(h))))
    launch_elu(lc, A, (size_t)T * s.cin, B);
    launch_conv_gemm(lc, B, s.res1, C, nullptr, nullptr, (int)T, (int)T, CE_STORE);
    launch_elu(lc, C, (size_t)T * (s.cin / 2), B);
    launch_conv_gemm(lc, B, s.res2, A, A, nullptr, (int)T, (int)T, CE_RES_SCALE);
    // ELU + downsampling conv (K = 2r, stride r) over the [T_out, r * C] view of the zero-padded activation
    const int64_t To = (T + s.ratio - 1) / s.ratio;
    launch_elu(lc, A, (size_t)T * s.cin, B);
    if (To * s.ratio > T) Q3_CUDA(cudaMemsetAsync(B + (size_t)T * s.cin, 0, sizeof(float) * (size_t)(To * s.ratio - T) * s.cin, stream_));
    launch_conv_gemm(lc, B, s.down, C, nullptr, nullptr, (int)To, (int)To, CE_STORE);
    std::swap(A, C);
    T = To;
  }
  launch_elu(lc, A, (size_t)T * conv_last_.cin, B);
  launch_conv_gemm(lc, B, conv_last_, C, nullptr, nullptr, (int)T, (int)T, CE_STORE);
  std::swap(A, C);
  // encoder transformer (:194-335): bidirectional, RoPE, LayerScale on both branches
  const int H = c.hidden_size, nh = c.num_attention_heads, nkv = c.num_key_value_heads, qkvw = (nh + 2 * nkv) * c.head_dim;
  for (const TLayer& l : tl_) {
    launch_layernorm(lc, A, (int)T, H, l.ln1_w, l.ln1_b, c.norm_eps, B);
    launch_conv_gemm(lc, B, l.qkv, C, nullptr, nullptr, (int)T, (int)T, CE_STORE);
    launch_codec_rope(lc, C, qkvw, (int)T, (int)T, nh + nkv, d_inv_freq_);
    launch_codec_attention_bidir(lc, C, qkvw, 1, (int)T, nh, nkv, B, nh * c.head_dim);
    launch_conv_gemm(lc, B, l.o, A, A, l.attn_scale, (int)T, (int)T, CE_RES_SCALE);
    launch_layernorm(lc, A, (int)T, H, l.ln2_w, l.ln2_b, c.norm_eps, B);
    launch_conv_gemm(lc, B, l.fc1, C, nullptr, nullptr, (int)T, (int)T, CE_GELU);
    launch_conv_gemm(lc, C, l.fc2, A, A, l.mlp_scale, (int)T, (int)T, CE_RES_SCALE);
  }
  // downsample (:339-358)
  const int64_t T2 = (T + c.compress - 1) / c.compress;
  if (T2 * c.compress > T) Q3_CUDA(cudaMemsetAsync(A + (size_t)T * H, 0, sizeof(float) * (size_t)(T2 * c.compress - T) * H, stream_));
  launch_conv_gemm(lc, A, downsample_, B, nullptr, nullptr, (int)T2, (int)T2, CE_STORE);
  Q3_CHECK(T2 == frames, Q3TTS_ERR_CUDA, "internal: encoder produced %lld frames, expected %d", (long long)T2, frames);
  if (h_latent) Q3_CUDA(cudaMemcpyAsync(h_latent, B, sizeof(float) * (size_t)T2 * H, cudaMemcpyDeviceToHost, stream_));
  // split residual vector quantiser (:424-460)
  const int D = c.vector_quantization_hidden_dimension;
  launch_conv_gemm(lc, B, proj_sem_, C, nullptr, nullptr, (int)T2, (int)T2, CE_STORE);
  launch_conv_gemm(lc, B, proj_ac_, C + (size_t)T2 * D, nullptr, nullptr, (int)T2, (int)T2, CE_STORE);
  launch_rvq_encode(lc, C, C + (size_t)T2 * D, d_books_, d_books_sq_, n_sem_, n_out_, D, c.codebook_size, (int)T2, d_codes_);
  Q3_CUDA(cudaMemcpyAsync(h_codes, d_codes_, sizeof(int32_t) * (size_t)n_out_ * (size_t)T2, cudaMemcpyDeviceToHost, stream_));
  Q3_CUDA(cudaStreamSynchronize(stream_));
  return frames;
}

}  // namespace q3

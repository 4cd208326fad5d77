// ICL reference-audio encoder on device: 24 kHz PCM -> 12.5 Hz codes [16][T].  Replaces Qwen3TTSAudioEncoder
// (Vocoder/Qwen3TTSAudioEncoder.swift:117-649): SEANet CNN (:120-190) -> bidirectional transformer (:194-335) -> stride-2
// downsample (:339-358) -> split residual vector quantiser, nearest-neighbour search (:362-460; EuclideanCodebook.encode,
// Vocoder/SpeechTokenizer.swift:511-519).  Activations are channels-last [T, C] (the reference transposes NCL <-> NLC around
// every conv); strided convs (kernel 2r, stride r, causal) run as 2-tap convs over the [T / r, r * C] view of their input.
#pragma once
#include <string>
#include <vector>

#include "codec.h"

namespace q3 {

struct AudioEncoderConfig {  // Qwen3TTSTokenizerEncoderConfig (Vocoder/SpeechTokenizer.swift:9-40)
  int audio_channels = 1, codebook_dim = 256, codebook_size = 2048, compress = 2, dilation_growth_rate = 2, hidden_size = 512;
  int intermediate_size = 2048, kernel_size = 7, last_kernel_size = 3, num_filters = 64, num_hidden_layers = 8, num_residual_layers = 1;
  int num_quantizers = 32, num_semantic_quantizers = 1, head_dim = 64, num_attention_heads = 8, num_key_value_heads = 8;
  int vector_quantization_hidden_dimension = 256, valid_num_quantizers = 16;
  float norm_eps = 1e-5f, rope_theta = 10000.0f;
  std::vector<int> upsampling_ratios{8, 6, 5, 4};
};

class AudioEncoder {
 public:
  // false when speech_tokenizer/model.safetensors carries no `encoder.*` tensors (the reference then has no usable encoder either)
  static bool present(const std::string& tokenizer_dir);
  AudioEncoder(const std::string& tokenizer_dir, cudaStream_t stream, LaunchCounter* counter);
  ~AudioEncoder();
  const AudioEncoderConfig& config() const { return cfg_; }
  int quantizers_out() const { return n_out_; }
  // frames produced for n_samples of audio (two rounds of ceil division: the CNN's strides, then the downsample)
  int frames_for(int64_t n_samples) const;
  // h_audio [n_samples] fp32 (host) -> h_codes [quantizers_out()][frames] int32 (host), row-major; returns frames.
  // h_latent (optional, host, [frames][hidden]) receives the quantiser's input (parity probe).
  int encode(const float* h_audio, int64_t n_samples, int32_t* h_codes, int capacity_frames, float* h_latent = nullptr);
  size_t device_bytes() const { return arena_.total() + ws_bytes_; }

 private:
  struct Stage { ConvW res1, res2, down; int ratio, cin, cout; };
  struct TLayer { ConvW qkv, o, fc1, fc2; const float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *attn_scale, *mlp_scale; };
  ConvW load_conv(const std::map<std::string, STensor>& t, const std::string& key, int cout, int cin, int k, int stride);
  ConvW load_linear(const std::map<std::string, STensor>& t, const std::vector<std::string>& keys, const std::vector<int>& outs, int cin, bool bias);
  const float* load_vec(const std::map<std::string, STensor>& t, const std::string& key, int n);
  void ensure_workspace(int64_t samples);
  LaunchCtx ctx() const { return LaunchCtx{stream_, counter_}; }

  AudioEncoderConfig cfg_;
  cudaStream_t stream_;
  LaunchCounter* counter_;
  DeviceArena arena_;
  ConvW conv0_, conv_last_, downsample_;
  std::vector<Stage> stages_;
  std::vector<TLayer> tl_;
  ConvW proj_sem_, proj_ac_;
  std::vector<const float*> books_;      // n_out_ x [codebook_size][D]
  std::vector<const float*> books_sq_;   // n_out_ x [codebook_size]  |e|^2
  const float** d_books_ = nullptr;
  const float** d_books_sq_ = nullptr;
  const float* d_inv_freq_ = nullptr;
  int n_out_ = 16, n_sem_ = 1;
  float* ws_[3] = {nullptr, nullptr, nullptr};
  int32_t* d_codes_ = nullptr;
  size_t ws_floats_ = 0, ws_bytes_ = 0;
  int64_t ws_samples_ = 0;
};

}  // namespace q3

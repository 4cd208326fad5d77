// ECAPA-TDNN speaker encoder on device: 24 kHz PCM -> speaker embedding [enc_dim].  Replaces SpeakerEncoder
// (SpeakerEncoder/SpeakerEncoder.swift:37-604): log-mel front end (reflect-padded STFT 1024 / hop 256, symmetric Hann, Slaney mel filterbank,
// :37-209), TimeDelayNet blocks with reflect padding (:234-257), SE-Res2Net blocks (:260-353), attentive statistics pooling (:355-397),
// final 1x1 conv (:496-524).  fp32 end to end; activations are channels-last [T, C] (the reference transposes NCL <-> NLC around every
// conv).  A per-voice, run-once operation: plain SIMT kernels, no tensor cores -- ~2 GFLOP for 3 s of audio.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "kernels.h"
#include "model.h"
#include "safetensors.h"

namespace q3 {

class SpeakerEncoderDev {
 public:
  // false when model.safetensors carries no `speaker_encoder.*` tensors (the reference then leaves speakerEncoder nil, Qwen3TTSPipeline.swift:156-169)
  static bool present(const std::string& model_dir);
  SpeakerEncoderDev(const std::string& model_dir, cudaStream_t stream, LaunchCounter* counter);
  ~SpeakerEncoderDev();
  int embedding_dim() const { return enc_dim_; }
  int frames_for(int64_t n_samples) const { return n_samples < 2 ? 0 : (int)(n_samples / 256 + 1); }  // (n + 2 * 512 - 1024) / 256 + 1
  // h_audio [n_samples] fp32 (host) -> h_embedding [embedding_dim()] (host).  h_mels (optional, host, [frames][128]): the log-mel input (parity probe).
  void extract(const float* h_audio, int64_t n_samples, float* h_embedding, float* h_mels = nullptr);
  size_t device_bytes() const { return arena_.total() + ws_bytes_; }

 private:
  struct Conv { const float* w = nullptr; const float* b = nullptr; int cout = 0, cin = 0, k = 1; };  // w: [k][cout][cin]
  struct SEBlock { Conv tdnn1, res[7], tdnn2, se1, se2; };
  Conv load_conv(const std::map<std::string, STensor>& t, const std::string& key);
  LaunchCtx ctx() const { return LaunchCtx{stream_, counter_}; }
  void ensure_workspace(int64_t n_samples, int frames);

  cudaStream_t stream_;
  LaunchCounter* counter_;
  DeviceArena arena_;
  Conv block0_, mfa_, asp_tdnn_, asp_conv_, fc_;
  SEBlock se_[3];
  const float *d_cos_ = nullptr, *d_sin_ = nullptr, *d_window_ = nullptr, *d_fb_ = nullptr;
  int ch_ = 0, mfa_ch_ = 0, enc_dim_ = 0;
  float* ws_ = nullptr;
  size_t ws_floats_ = 0, ws_bytes_ = 0;
};

}  // namespace q3

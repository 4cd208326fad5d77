// mmap-backed safetensors reader.  Replaces `MLX.loadArrays(url:)` (Qwen3TTSPipeline.swift:142,
// Vocoder/AudioDecoder.swift:141): 8-byte little-endian header length, JSON header, raw little-endian data.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <map>
#include <string>
#include <vector>

#include "json.h"

namespace q3 {

struct STensor {
  std::string dtype;  // "F32" "F16" "BF16" "U32" "I32" "U8" ...
  std::vector<int64_t> shape;
  const uint8_t* data = nullptr;
  size_t nbytes = 0;
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : shape) n *= d;
    return n;
  }
  bool is_float() const { return dtype == "F32" || dtype == "F16" || dtype == "BF16"; }
  int q3_dtype() const {
    if (dtype == "F32") return Q3TTS_F32;
    if (dtype == "F16") return Q3TTS_F16;
    if (dtype == "BF16") return Q3TTS_BF16;
    fail(Q3TTS_ERR_BAD_WEIGHTS, "tensor dtype %s is not a float type", dtype.c_str());
  }
};

class SafeTensors {
 public:
  explicit SafeTensors(const std::string& path, q3tts_status missing = Q3TTS_ERR_FILE_NOT_FOUND) {
    fd_ = open(path.c_str(), O_RDONLY);
    if (fd_ < 0) fail(missing, "Required file not found: %s", path.c_str());
    struct stat st;
    if (fstat(fd_, &st) != 0 || st.st_size < 8) {
      close(fd_);
      fail(Q3TTS_ERR_BAD_WEIGHTS, "cannot stat %s", path.c_str());
    }
    size_ = (size_t)st.st_size;
    base_ = (const uint8_t*)mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
    if (base_ == MAP_FAILED) {
      close(fd_);
      base_ = nullptr;
      fail(Q3TTS_ERR_BAD_WEIGHTS, "mmap failed for %s", path.c_str());
    }
    uint64_t hlen = 0;
    memcpy(&hlen, base_, 8);
    if (hlen + 8 > size_) fail(Q3TTS_ERR_BAD_WEIGHTS, "corrupt safetensors header in %s", path.c_str());
    Json h = JsonParser((const char*)base_ + 8, (size_t)hlen).parse();
    const uint8_t* data0 = base_ + 8 + hlen;
    for (auto& kv : h.obj) {
      if (kv.first == "__metadata__") continue;
      STensor t;
      t.dtype = kv.second.at("dtype").str;
      for (auto& d : kv.second.at("shape").arr) t.shape.push_back((int64_t)llround(d.num));
      auto& off = kv.second.at("data_offsets").arr;
      size_t b = (size_t)llround(off.at(0).num), e = (size_t)llround(off.at(1).num);
      if (e < b || 8 + hlen + e > size_) fail(Q3TTS_ERR_BAD_WEIGHTS, "tensor %s out of file bounds", kv.first.c_str());
      t.data = data0 + b;
      t.nbytes = e - b;
      tensors_[kv.first] = t;
    }
  }
  ~SafeTensors() {
    if (base_) munmap((void*)base_, size_);
    if (fd_ >= 0) close(fd_);
  }
  SafeTensors(const SafeTensors&) = delete;
  SafeTensors& operator=(const SafeTensors&) = delete;
  const std::map<std::string, STensor>& tensors() const { return tensors_; }

 private:
  int fd_ = -1;
  const uint8_t* base_ = nullptr;
  size_t size_ = 0;
  std::map<std::string, STensor> tensors_;
};

}  // namespace q3

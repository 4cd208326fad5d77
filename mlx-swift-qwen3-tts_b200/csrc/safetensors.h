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
    try {
      parse(path);
    } catch (...) {  // a throwing constructor never runs the destructor: release the mapping and the descriptor here
      unmap();
      throw;
    }
  }
  ~SafeTensors() { unmap(); }
  SafeTensors(const SafeTensors&) = delete;
  SafeTensors& operator=(const SafeTensors&) = delete;
  const std::map<std::string, STensor>& tensors() const { return tensors_; }

 private:
  static size_t elem_size(const std::string& dt) {
    if (dt == "F64" || dt == "I64" || dt == "U64") return 8;
    if (dt == "F32" || dt == "I32" || dt == "U32") return 4;
    if (dt == "F16" || dt == "BF16" || dt == "I16" || dt == "U16") return 2;
    if (dt == "I8" || dt == "U8" || dt == "BOOL" || dt == "F8_E4M3" || dt == "F8_E5M2") return 1;
    return 0;
  }
  void unmap() {
    if (base_) munmap((void*)base_, size_);
    if (fd_ >= 0) close(fd_);
    base_ = nullptr;
    fd_ = -1;
  }
  // The file is untrusted input: every offset, dimension and byte count is checked against the mapping before a tensor is
  // handed out, so the loaders may size device buffers from (shape, dtype) and copy `nbytes` without reading past the file.
  void parse(const std::string& path) {
    struct stat st;
    if (fstat(fd_, &st) != 0 || st.st_size < 8) fail(Q3TTS_ERR_BAD_WEIGHTS, "cannot stat %s", path.c_str());
    size_ = (size_t)st.st_size;
    void* m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
    if (m == MAP_FAILED) fail(Q3TTS_ERR_BAD_WEIGHTS, "mmap failed for %s", path.c_str());
    base_ = (const uint8_t*)m;
    uint64_t hlen = 0;
    memcpy(&hlen, base_, 8);
    if (hlen > (uint64_t)(size_ - 8)) fail(Q3TTS_ERR_BAD_WEIGHTS, "corrupt safetensors header in %s", path.c_str());
    Json h = JsonParser((const char*)base_ + 8, (size_t)hlen).parse();
    const uint8_t* data0 = base_ + 8 + hlen;
    const size_t data_bytes = size_ - 8 - (size_t)hlen;
    for (auto& kv : h.obj) {
      if (kv.first == "__metadata__") continue;
      STensor t;
      t.dtype = kv.second.at("dtype").str;
      const size_t esz = elem_size(t.dtype);
      if (esz == 0) fail(Q3TTS_ERR_BAD_WEIGHTS, "tensor %s has unknown dtype %s", kv.first.c_str(), t.dtype.c_str());
      uint64_t numel = 1;
      for (auto& d : kv.second.at("shape").arr) {
        if (!(d.num >= 0.0) || d.num > 9.0e15) fail(Q3TTS_ERR_BAD_WEIGHTS, "tensor %s has an invalid dimension", kv.first.c_str());
        const int64_t dim = (int64_t)llround(d.num);
        if (dim != 0 && numel > (UINT64_MAX / 16) / (uint64_t)dim) fail(Q3TTS_ERR_BAD_WEIGHTS, "tensor %s: element count overflows", kv.first.c_str());
        numel *= (uint64_t)dim;
        t.shape.push_back(dim);
      }
      auto& off = kv.second.at("data_offsets").arr;
      if (off.size() != 2 || !(off[0].num >= 0.0) || !(off[1].num >= off[0].num) || off[1].num > (double)data_bytes)
        fail(Q3TTS_ERR_BAD_WEIGHTS, "tensor %s out of file bounds", kv.first.c_str());
      const size_t b = (size_t)llround(off[0].num), e = (size_t)llround(off[1].num);
      if (e < b || e > data_bytes) fail(Q3TTS_ERR_BAD_WEIGHTS, "tensor %s out of file bounds", kv.first.c_str());
      if ((uint64_t)(e - b) != numel * (uint64_t)esz)
        fail(Q3TTS_ERR_BAD_WEIGHTS, "tensor %s: %zu data bytes do not match its shape and dtype %s", kv.first.c_str(), e - b, t.dtype.c_str());
      t.data = data0 + b;
      t.nbytes = e - b;
      tensors_[kv.first] = t;
    }
  }
  int fd_ = -1;
  const uint8_t* base_ = nullptr;
  size_t size_ = 0;
  std::map<std::string, STensor> tensors_;
};

}  // namespace q3

// Minimal JSON reader (objects, arrays, strings, numbers, true/false/null) — enough for config.json and the
// safetensors header.  Replaces Foundation's JSONDecoder on the reference side (Qwen3TTSPipeline.swift:131-132).
#pragma once
#include <cmath>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.h"

namespace q3 {

struct Json {
  enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;  // insertion order kept

  const Json* find(const std::string& k) const {
    if (type != Obj) return nullptr;
    for (auto& kv : obj)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
  bool has(const std::string& k) const { return find(k) != nullptr; }
  const Json& at(const std::string& k) const {
    const Json* j = find(k);
    if (!j) fail(Q3TTS_ERR_BAD_CONFIG, "missing JSON key '%s'", k.c_str());
    return *j;
  }
  double number(const std::string& k) const {
    const Json& j = at(k);
    if (j.type != Num) fail(Q3TTS_ERR_BAD_CONFIG, "JSON key '%s' is not a number", k.c_str());
    return j.num;
  }
  double number_or(const std::string& k, double d) const {
    const Json* j = find(k);
    return (j && j->type == Num) ? j->num : d;
  }
  int integer(const std::string& k) const { return (int)llround(number(k)); }
  int integer_or(const std::string& k, int d) const { return (int)llround(number_or(k, d)); }
  bool bool_or(const std::string& k, bool d) const {
    const Json* j = find(k);
    return (j && j->type == Bool) ? j->b : d;
  }
  std::string string_or(const std::string& k, const std::string& d) const {
    const Json* j = find(k);
    return (j && j->type == Str) ? j->str : d;
  }
  std::vector<int> int_array_or(const std::string& k, std::vector<int> d) const {
    const Json* j = find(k);
    if (!j || j->type != Arr) return d;
    std::vector<int> r;
    for (auto& e : j->arr) r.push_back((int)llround(e.num));
    return r;
  }
};

class JsonParser {
 public:
  JsonParser(const char* p, size_t n) : p_(p), e_(p + n) {}
  Json parse() {
    Json j = value();
    ws();
    return j;
  }

 private:
  const char *p_, *e_;
  void ws() {
    while (p_ < e_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_;
  }
  [[noreturn]] void bad(const char* what) { fail(Q3TTS_ERR_BAD_CONFIG, "JSON parse error: %s", what); }
  Json value() {
    ws();
    if (p_ >= e_) bad("unexpected end");
    Json j;
    char c = *p_;
    if (c == '{') {
      j.type = Json::Obj;
      ++p_;
      ws();
      if (p_ < e_ && *p_ == '}') { ++p_; return j; }
      for (;;) {
        ws();
        std::string k = string();
        ws();
        if (p_ >= e_ || *p_ != ':') bad("expected ':'");
        ++p_;
        j.obj.emplace_back(std::move(k), value());
        ws();
        if (p_ < e_ && *p_ == ',') { ++p_; continue; }
        if (p_ < e_ && *p_ == '}') { ++p_; break; }
        bad("expected ',' or '}'");
      }
    } else if (c == '[') {
      j.type = Json::Arr;
      ++p_;
      ws();
      if (p_ < e_ && *p_ == ']') { ++p_; return j; }
      for (;;) {
        j.arr.push_back(value());
        ws();
        if (p_ < e_ && *p_ == ',') { ++p_; continue; }
        if (p_ < e_ && *p_ == ']') { ++p_; break; }
        bad("expected ',' or ']'");
      }
    } else if (c == '"') {
      j.type = Json::Str;
      j.str = string();
    } else if (c == 't' && e_ - p_ >= 4 && !strncmp(p_, "true", 4)) {
      j.type = Json::Bool; j.b = true; p_ += 4;
    } else if (c == 'f' && e_ - p_ >= 5 && !strncmp(p_, "false", 5)) {
      j.type = Json::Bool; j.b = false; p_ += 5;
    } else if (c == 'n' && e_ - p_ >= 4 && !strncmp(p_, "null", 4)) {
      p_ += 4;
    } else if (c == 'N' && e_ - p_ >= 3 && !strncmp(p_, "NaN", 3)) {
      j.type = Json::Num; j.num = NAN; p_ += 3;
    } else {
      char* end = nullptr;
      std::string tmp(p_, (size_t)std::min<ptrdiff_t>(e_ - p_, 64));
      j.num = strtod(tmp.c_str(), &end);
      if (end == tmp.c_str()) bad("bad number");
      j.type = Json::Num;
      p_ += end - tmp.c_str();
    }
    return j;
  }
  std::string string() {
    if (p_ >= e_ || *p_ != '"') bad("expected string");
    ++p_;
    std::string s;
    while (p_ < e_ && *p_ != '"') {
      if (*p_ == '\\' && p_ + 1 < e_) {
        ++p_;
        switch (*p_) {
          case 'n': s += '\n'; break;
          case 't': s += '\t'; break;
          case 'r': s += '\r'; break;
          case 'b': s += '\b'; break;
          case 'f': s += '\f'; break;
          case 'u': {
            if (e_ - p_ < 5) bad("bad \\u escape");
            unsigned cp = (unsigned)strtoul(std::string(p_ + 1, 4).c_str(), nullptr, 16);
            p_ += 4;
            if (cp < 0x80) s += (char)cp;
            else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
            else { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: s += *p_;
        }
        ++p_;
      } else {
        s += *p_++;
      }
    }
    if (p_ >= e_) bad("unterminated string");
    ++p_;
    return s;
  }
};

inline Json parse_json_file(const std::string& path, q3tts_status missing = Q3TTS_ERR_FILE_NOT_FOUND) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) fail(missing, "Required file not found: %s", path.c_str());
  std::string buf;
  char tmp[65536];
  size_t n;
  while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.append(tmp, n);
  fclose(f);
  return JsonParser(buf.data(), buf.size()).parse();
}

}  // namespace q3

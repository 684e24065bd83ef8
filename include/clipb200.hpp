// C++17 host-side mirror of the `open_clip_inference` crate's public API (RuurdBijlsma/clip-embedder-rs) over the
// C ABI of clipb200.h — the compiled-language counterpart of the Python mirror in clip_embedder_rs_b200/*.py, for
// callers that would otherwise link the Rust crate.  Header-only; link with -lclipb200.
//
//   reference (Rust)                                           here
//   ---------------------------------------------------------  -------------------------------------------------
//   ClipError (src/error.rs:9-41)                              clipb200::ClipError { kind(), what() }  same messages
//   ModelConfig / OpenClipConfig::from_file (src/config.rs)    clipb200::ModelConfig / OpenClipConfig ::from_file
//   model_manager::{MODEL_FILES, verify_model_dir,             clipb200::model_manager::{MODEL_FILES, verify_model_dir,
//       get_default_base_folder} (src/model_manager.rs:8-68)       get_default_base_folder}
//   OnnxSession::{new, has_input, find_input} (src/onnx.rs)    clipb200::OnnxSession
//   VisionEmbedder::{from_local_dir, from_local_id, duplicate, clipb200::VisionEmbedder (images are RGB8 views: what
//       embed_image(s), preprocess(_batch)} (src/vision.rs)        `DynamicImage::to_rgb8` yields, src/vision.rs:171)
//   TextEmbedder::{from_local_dir, from_local_id, duplicate,   clipb200::TextEmbedder (the `tokenizers` crate has no
//       tokenize, embed_text(s)} (src/text.rs)                     C++ port: ids come from a caller-supplied encoder)
//   Clip::{from_local_dir, from_local_id, duplicate,           clipb200::Clip
//       get_model_config, compare, classify, rank_images,
//       softmax, sigmoid} (src/clip.rs)
//
// `from_hf` is not mirrored (network).  `with_execution_providers` is accepted as a CUDA device index: the only
// provider this library has is the sm_100a engine, and there is no CPU fallback.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <sys/stat.h>

#include <algorithm>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "clipb200.h"

namespace clipb200 {

// ------------------------------------------------------------------------------------------------ errors
enum class ErrorKind { Io, Json, Ort, Image, Tokenizer, Config, Inference, Shape, ModelFolderNotFound, MissingModelFile,
                       LockPoison };

class ClipError : public std::runtime_error {  // src/error.rs:9-41 (same Display strings)
 public:
  ClipError(ErrorKind kind, const std::string& msg) : std::runtime_error(msg), kind_(kind) {}
  ErrorKind kind() const { return kind_; }
  static ClipError Io(const std::string& m) { return {ErrorKind::Io, "IO error: " + m}; }
  static ClipError Json(const std::string& m) { return {ErrorKind::Json, "JSON error: " + m}; }
  static ClipError Ort(const std::string& m) { return {ErrorKind::Ort, "ONNX Runtime Error: " + m}; }
  static ClipError Tokenizer(const std::string& m) { return {ErrorKind::Tokenizer, "Tokenization error: " + m}; }
  static ClipError Config(const std::string& m) { return {ErrorKind::Config, "Configuration error: " + m}; }
  static ClipError Inference(const std::string& m) { return {ErrorKind::Inference, "Inference error: " + m}; }
  static ClipError Shape(const std::string& m) { return {ErrorKind::Shape, "Shape error: " + m}; }
  static ClipError ModelFolderNotFound(const std::string& dir) {
    return {ErrorKind::ModelFolderNotFound,
            "Model folder not found, generate it with `uv run pull_onnx.py -h`. '" + dir + "'"};
  }
  static ClipError MissingModelFile(const std::string& dir, const std::string& file) {
    return {ErrorKind::MissingModelFile, "Missing model file '" + file + "' in folder '" + dir + "'"};
  }

 private:
  ErrorKind kind_;
};

namespace detail {

// Minimal JSON reader for the two config files (objects, arrays, strings, numbers, true/false/null).
struct Json {
  enum Type { Null, Bool, Number, String, Array, Object } type = Null;
  bool b = false;
  double num = 0.0;
  std::string str;
  std::vector<Json> arr;
  std::map<std::string, Json> obj;
  const Json* get(const std::string& k) const {
    if (type != Object) return nullptr;
    auto it = obj.find(k);
    return it == obj.end() || it->second.type == Null ? nullptr : &it->second;
  }
};

class JsonParser {
 public:
  explicit JsonParser(const std::string& s) : s_(s) {}
  Json parse() {
    Json v = value();
    ws();
    if (p_ != s_.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string& s_;
  size_t p_ = 0;
  [[noreturn]] void fail(const std::string& m) const {
    throw ClipError::Json(m + " at offset " + std::to_string(p_));
  }
  void ws() { while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\n' || s_[p_] == '\t' || s_[p_] == '\r')) ++p_; }
  Json value() {
    ws();
    if (p_ >= s_.size()) fail("EOF while parsing a value");
    const char c = s_[p_];
    Json v;
    if (c == '{') {
      v.type = Json::Object;
      ++p_;
      ws();
      if (p_ < s_.size() && s_[p_] == '}') { ++p_; return v; }
      for (;;) {
        ws();
        if (p_ >= s_.size() || s_[p_] != '"') fail("key must be a string");
        std::string k = string();
        ws();
        if (p_ >= s_.size() || s_[p_] != ':') fail("expected `:`");
        ++p_;
        v.obj[k] = value();
        ws();
        if (p_ < s_.size() && s_[p_] == ',') { ++p_; continue; }
        if (p_ < s_.size() && s_[p_] == '}') { ++p_; return v; }
        fail("expected `,` or `}`");
      }
    }
    if (c == '[') {
      v.type = Json::Array;
      ++p_;
      ws();
      if (p_ < s_.size() && s_[p_] == ']') { ++p_; return v; }
      for (;;) {
        v.arr.push_back(value());
        ws();
        if (p_ < s_.size() && s_[p_] == ',') { ++p_; continue; }
        if (p_ < s_.size() && s_[p_] == ']') { ++p_; return v; }
        fail("expected `,` or `]`");
      }
    }
    if (c == '"') { v.type = Json::String; v.str = string(); return v; }
    if (s_.compare(p_, 4, "true") == 0) { p_ += 4; v.type = Json::Bool; v.b = true; return v; }
    if (s_.compare(p_, 5, "false") == 0) { p_ += 5; v.type = Json::Bool; v.b = false; return v; }
    if (s_.compare(p_, 4, "null") == 0) { p_ += 4; return v; }
    char* end = nullptr;
    v.num = strtod(s_.c_str() + p_, &end);
    if (end == s_.c_str() + p_) fail("expected value");
    p_ = static_cast<size_t>(end - s_.c_str());
    v.type = Json::Number;
    return v;
  }
  std::string string() {
    std::string out;
    ++p_;
    while (p_ < s_.size() && s_[p_] != '"') {
      char c = s_[p_++];
      if (c == '\\' && p_ < s_.size()) {
        const char e = s_[p_++];
        switch (e) {
          case 'n': out.push_back('\n'); break;
          case 't': out.push_back('\t'); break;
          case 'r': out.push_back('\r'); break;
          case 'b': out.push_back('\b'); break;
          case 'f': out.push_back('\f'); break;
          case 'u': {
            if (p_ + 4 > s_.size()) fail("bad \\u escape");
            const unsigned cp = static_cast<unsigned>(strtoul(s_.substr(p_, 4).c_str(), nullptr, 16));
            p_ += 4;
            if (cp < 0x80) out.push_back(static_cast<char>(cp));
            else if (cp < 0x800) { out.push_back(static_cast<char>(0xC0 | (cp >> 6))); out.push_back(static_cast<char>(0x80 | (cp & 0x3F))); }
            else { out.push_back(static_cast<char>(0xE0 | (cp >> 12))); out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F))); out.push_back(static_cast<char>(0x80 | (cp & 0x3F))); }
            break;
          }
          default: out.push_back(e);
        }
      } else {
        out.push_back(c);
      }
    }
    if (p_ >= s_.size()) fail("EOF while parsing a string");
    ++p_;
    return out;
  }
};

inline std::string read_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw ClipError::Io("No such file or directory (os error 2): " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}
inline const Json& need(const Json& o, const char* key, Json::Type t) {
  const Json* v = o.get(key);
  if (v == nullptr) throw ClipError::Json(std::string("missing field `") + key + "`");
  if (v->type != t) throw ClipError::Json(std::string("invalid type for field `") + key + "`");
  return *v;
}
inline bool is_dir(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }
inline bool is_file(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode); }
inline bool exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0; }
inline std::string join(const std::string& a, const std::string& b) {
  return a.empty() || a.back() == '/' ? a + b : a + "/" + b;
}

}  // namespace detail

// ------------------------------------------------------------------------------------------------ configs
struct ModelConfig {  // src/config.rs:7-23 (model_config.json written by pull_onnx.py:143-150)
  bool tokenizer_needs_lowercase = false;
  std::optional<std::string> activation_function;
  std::optional<float> logit_scale, logit_bias;
  std::optional<uint32_t> pad_id;
  static ModelConfig from_file(const std::string& path) {
    const detail::Json j = detail::JsonParser(detail::read_file(path)).parse();
    if (j.type != detail::Json::Object) throw ClipError::Json("invalid type: expected a map");
    ModelConfig c;
    if (const detail::Json* v = j.get("tokenizer_needs_lowercase")) c.tokenizer_needs_lowercase = v->b;
    if (const detail::Json* v = j.get("activation_function")) c.activation_function = v->str;
    if (const detail::Json* v = j.get("logit_scale")) c.logit_scale = static_cast<float>(v->num);
    if (const detail::Json* v = j.get("logit_bias")) c.logit_bias = static_cast<float>(v->num);
    if (const detail::Json* v = j.get("pad_id")) c.pad_id = static_cast<uint32_t>(v->num);
    return c;
  }
};

struct VisionCfg { uint32_t image_size = 0; std::optional<size_t> layers, width; };
struct TextCfg { size_t context_length = 0; std::optional<std::string> hf_tokenizer_name; };
struct ModelCfg { size_t embed_dim = 0; VisionCfg vision_cfg; TextCfg text_cfg; };
struct PreprocessCfg {
  float mean[3] = {0, 0, 0}, std[3] = {1, 1, 1};
  std::string interpolation = "bicubic";  // serde default (src/config.rs:59-61)
  std::string resize_mode = "shortest";   // serde default (src/config.rs:62-64)
};
struct OpenClipConfig {  // src/config.rs:24-71
  ModelCfg model_cfg;
  PreprocessCfg preprocess_cfg;
  static OpenClipConfig from_file(const std::string& path) {
    using detail::Json;
    const Json j = detail::JsonParser(detail::read_file(path)).parse();
    if (j.type != Json::Object) throw ClipError::Json("invalid type: expected a map");
    OpenClipConfig c;
    const Json& m = detail::need(j, "model_cfg", Json::Object);
    c.model_cfg.embed_dim = static_cast<size_t>(detail::need(m, "embed_dim", Json::Number).num);
    const Json& v = detail::need(m, "vision_cfg", Json::Object);
    c.model_cfg.vision_cfg.image_size = static_cast<uint32_t>(detail::need(v, "image_size", Json::Number).num);
    if (const Json* x = v.get("layers")) if (x->type == Json::Number) c.model_cfg.vision_cfg.layers = static_cast<size_t>(x->num);
    if (const Json* x = v.get("width")) if (x->type == Json::Number) c.model_cfg.vision_cfg.width = static_cast<size_t>(x->num);
    const Json& t = detail::need(m, "text_cfg", Json::Object);
    c.model_cfg.text_cfg.context_length = static_cast<size_t>(detail::need(t, "context_length", Json::Number).num);
    if (const Json* x = t.get("hf_tokenizer_name")) c.model_cfg.text_cfg.hf_tokenizer_name = x->str;
    const Json& p = detail::need(j, "preprocess_cfg", Json::Object);
    const Json& mean = detail::need(p, "mean", Json::Array);
    const Json& sd = detail::need(p, "std", Json::Array);
    if (mean.arr.size() != 3 || sd.arr.size() != 3) throw ClipError::Json("invalid length, expected an array of length 3");
    for (int i = 0; i < 3; ++i) {
      c.preprocess_cfg.mean[i] = static_cast<float>(mean.arr[i].num);
      c.preprocess_cfg.std[i] = static_cast<float>(sd.arr[i].num);
    }
    if (const Json* x = p.get("interpolation")) c.preprocess_cfg.interpolation = x->str;
    if (const Json* x = p.get("resize_mode")) c.preprocess_cfg.resize_mode = x->str;
    return c;
  }
};

namespace model_manager {  // src/model_manager.rs:8-68
static const char* const MODEL_FILES[] = {"model_config.json", "open_clip_config.json", "special_tokens_map.json",
                                          "text.onnx", "tokenizer.json", "tokenizer_config.json", "visual.onnx",
                                          "text.onnx.data", "visual.onnx.data"};
inline std::string get_default_base_folder() {
  const char* home = getenv("HOME");
  return home != nullptr && *home ? detail::join(home, ".cache/open_clip_rs") : std::string(".open_clip_cache");
}
inline void verify_model_dir(const std::string& model_dir) {
  if (!detail::exists(model_dir)) throw ClipError::ModelFolderNotFound(model_dir);
  for (const char* f : MODEL_FILES)
    if (!detail::is_file(detail::join(model_dir, f))) throw ClipError::MissingModelFile(model_dir, f);
}
}  // namespace model_manager

// ------------------------------------------------------------------------------------------------ session
class OnnxSession {  // src/onnx.rs:8-47: the ort::Session becomes an engine handle
 public:
  OnnxSession(const std::string& path, int cuda_device = 0) : path_(path), device_(cuda_device) {
    clipb200_engine* h = nullptr;
    if (clipb200_engine_create(path.c_str(), cuda_device, nullptr, &h) != CLIPB200_OK) throw ClipError::Ort(clipb200_last_error());
    handle_.reset(h, [](clipb200_engine* e) { clipb200_engine_destroy(e); });
    mutex_ = std::make_shared<std::mutex>();
  }
  std::vector<std::string> input_names() const {
    std::vector<std::string> names;
    for (int i = 0, n = clipb200_engine_num_inputs(handle_.get()); i < n; ++i) names.push_back(clipb200_engine_input_name(handle_.get(), i));
    return names;
  }
  bool has_input(const std::string& name) const {  // onnx.rs:32-35
    const auto names = input_names();
    return std::find(names.begin(), names.end(), name) != names.end();
  }
  std::optional<std::string> find_input(const std::vector<std::string>& possibilities) const {  // onnx.rs:38-46
    for (const std::string& p : possibilities) if (has_input(p)) return p;
    return std::nullopt;
  }
  clipb200_engine* handle() const { return handle_.get(); }
  std::mutex& write_lock() const { return *mutex_; }  // the reference serialises runs with RwLock::write
  int64_t embed_dim() const { return clipb200_engine_embed_dim(handle_.get()); }
  const std::string& path() const { return path_; }
  int device() const { return device_; }
  static void check(int rc) { if (rc != CLIPB200_OK) throw ClipError::Ort(clipb200_last_error()); }

 private:
  std::string path_;
  int device_ = 0;
  std::shared_ptr<clipb200_engine> handle_;
  std::shared_ptr<std::mutex> mutex_;
};

// What `DynamicImage::to_rgb8()` yields (src/vision.rs:171): packed RGB8, row-major, no padding.  Borrowed.
struct RgbImage {
  const uint8_t* data = nullptr;
  int32_t width = 0, height = 0;
};

// ------------------------------------------------------------------------------------------------ vision
class VisionEmbedder {  // src/vision.rs
 public:
  OnnxSession session;
  OpenClipConfig config;
  std::string model_dir;
  std::string input_name;

  static VisionEmbedder from_local_dir(const std::string& model_dir, int cuda_device = 0) {  // vision.rs:58-84
    model_manager::verify_model_dir(model_dir);
    OpenClipConfig config = OpenClipConfig::from_file(detail::join(model_dir, "open_clip_config.json"));
    OnnxSession session(detail::join(model_dir, "visual.onnx"), cuda_device);
    auto name = session.find_input({"pixel_values", "input"});
    if (!name) throw ClipError::Config("Could not find vision input node");  // vision.rs:75
    return VisionEmbedder(std::move(session), std::move(config), model_dir, *name);
  }
  static VisionEmbedder from_local_id(const std::string& model_id, const std::string& base_folder = "", int cuda_device = 0) {
    return from_local_dir(detail::join(base_folder.empty() ? model_manager::get_default_base_folder() : base_folder, model_id), cuda_device);
  }
  VisionEmbedder duplicate() const { return from_local_dir(model_dir, session.device()); }  // vision.rs:87-91

  std::vector<float> embed_image(const RgbImage& image) const { return embed_images({image}); }  // vision.rs:94-98
  // -> row-major [images.size(), embed_dim], rows L2-normalised (vision.rs:102-117); resize + normalise + tower on the GPU
  std::vector<float> embed_images(const std::vector<RgbImage>& images) const {
    if (images.empty()) throw ClipError::Inference("Empty batch");  // vision.rs:121-123
    std::vector<const uint8_t*> ptrs;
    std::vector<int32_t> ws, hs;
    for (const RgbImage& im : images) {
      if (im.data == nullptr || im.width <= 0 || im.height <= 0) throw ClipError::Shape("image must be a non-empty RGB8 buffer");
      ptrs.push_back(im.data);
      ws.push_back(im.width);
      hs.push_back(im.height);
    }
    std::vector<float> out(images.size() * static_cast<size_t>(session.embed_dim()));
    const clipb200_preproc pp = preproc();
    std::lock_guard<std::mutex> guard(session.write_lock());
    OnnxSession::check(clipb200_vision_embed_rgb8_var(session.handle(), ptrs.data(), ws.data(), hs.data(),
                                                      static_cast<int64_t>(images.size()), &pp, out.data()));
    return out;
  }
  // `preprocess_batch` (vision.rs:120-135): -> f32 [B, 3, S, S]
  std::vector<float> preprocess_batch(const std::vector<RgbImage>& images) const {
    if (images.empty()) throw ClipError::Inference("Empty batch");
    const int32_t S = static_cast<int32_t>(config.model_cfg.vision_cfg.image_size);
    const size_t plane = static_cast<size_t>(S) * S * 3;
    std::vector<float> out(images.size() * plane);
    std::vector<uint8_t> resized(plane);
    const clipb200_preproc pp = preproc();
    std::lock_guard<std::mutex> guard(session.write_lock());
    for (size_t i = 0; i < images.size(); ++i) {
      const RgbImage& im = images[i];
      const uint8_t* src = im.data;
      if (im.width != S || im.height != S) {
        OnnxSession::check(clipb200_resize_rgb8(session.handle(), im.data, im.width, im.height, &pp, resized.data()));
        src = resized.data();
      }
      OnnxSession::check(clipb200_preprocess_rgb8(session.handle(), src, 1, S, S, &pp, out.data() + i * plane));
    }
    return out;
  }
  std::vector<float> preprocess(const RgbImage& image) const { return preprocess_batch({image}); }  // vision.rs:138-140

 private:
  VisionEmbedder(OnnxSession s, OpenClipConfig c, std::string dir, std::string in)
      : session(std::move(s)), config(std::move(c)), model_dir(std::move(dir)), input_name(std::move(in)) {}
  clipb200_preproc preproc() const {
    clipb200_preproc pp;
    for (int i = 0; i < 3; ++i) { pp.mean[i] = config.preprocess_cfg.mean[i]; pp.std[i] = config.preprocess_cfg.std[i]; }
    const std::string& ip = config.preprocess_cfg.interpolation;  // vision.rs:176-180
    pp.interpolation = ip == "bicubic" ? 0 : (ip == "bilinear" ? 1 : 2);
    pp.resize_mode = config.preprocess_cfg.resize_mode == "squash" ? 1 : 0;  // vision.rs:184-192
    return pp;
  }
};

// ------------------------------------------------------------------------------------------------ text
// Encoder for one string: ids WITH special tokens, at most `context_length` of them (what tokenizers'
// `encode(text, add_special_tokens = true)` returns under the truncation params of src/text.rs:76-85).
using EncodeFn = std::function<std::vector<uint32_t>(const std::string&)>;

class TextEmbedder {  // src/text.rs
 public:
  OnnxSession session;
  OpenClipConfig config;
  ModelConfig model_config;
  std::string model_dir;
  std::string id_name;
  std::optional<std::string> mask_name;
  EncodeFn encoder;  // must be set before tokenize / embed_text(s); embed_ids needs none

  static TextEmbedder from_local_dir(const std::string& model_dir, int cuda_device = 0) {  // text.rs:55-101
    model_manager::verify_model_dir(model_dir);
    OpenClipConfig config = OpenClipConfig::from_file(detail::join(model_dir, "open_clip_config.json"));
    ModelConfig model_config = ModelConfig::from_file(detail::join(model_dir, "model_config.json"));
    OnnxSession session(detail::join(model_dir, "text.onnx"), cuda_device);
    auto id = session.find_input({"input_ids"});
    if (!id) throw ClipError::Config("Could not find text input node");  // text.rs:89
    TextEmbedder t(std::move(session), std::move(config), std::move(model_config), model_dir, *id);
    t.mask_name = t.session.find_input({"attention_mask"});  // text.rs:90
    return t;
  }
  static TextEmbedder from_local_id(const std::string& model_id, const std::string& base_folder = "", int cuda_device = 0) {
    return from_local_dir(detail::join(base_folder.empty() ? model_manager::get_default_base_folder() : base_folder, model_id), cuda_device);
  }
  TextEmbedder duplicate() const {  // text.rs:104-108
    TextEmbedder t = from_local_dir(model_dir, session.device());
    t.encoder = encoder;
    return t;
  }

  size_t context_length() const { return config.model_cfg.text_cfg.context_length; }
  // text.rs:111-139: optional lowercase, encode with specials, right-pad to ctx with pad_id -> (ids, mask) [B, ctx]
  std::pair<std::vector<int64_t>, std::vector<int64_t>> tokenize(const std::vector<std::string>& texts) const {
    if (!encoder) throw ClipError::Tokenizer("no encoder set (TextEmbedder::encoder)");
    const size_t ctx = context_length();
    const int64_t pad = static_cast<int64_t>(model_config.pad_id.value_or(0));
    std::vector<int64_t> ids(texts.size() * ctx, pad), mask(texts.size() * ctx, 0);
    for (size_t i = 0; i < texts.size(); ++i) {
      std::string t = texts[i];
      if (model_config.tokenizer_needs_lowercase)  // Rust's to_lowercase is Unicode-aware; ASCII here
        for (char& c : t) if (c >= 'A' && c <= 'Z') c = static_cast<char>(c - 'A' + 'a');
      const std::vector<uint32_t> enc = encoder(t);
      if (enc.size() > ctx) throw ClipError::Tokenizer("encoder returned more than context_length ids");
      for (size_t k = 0; k < enc.size(); ++k) { ids[i * ctx + k] = enc[k]; mask[i * ctx + k] = 1; }
    }
    return {std::move(ids), std::move(mask)};
  }
  std::vector<float> embed_text(const std::string& text) const { return embed_texts({text}); }  // text.rs:142-146
  std::vector<float> embed_texts(const std::vector<std::string>& texts) const {                 // text.rs:150-169
    auto tm = tokenize(texts);
    return embed_ids(tm.first, mask_name ? &tm.second : nullptr, texts.size());
  }
  // the `session.run(inputs![input_ids (, attention_mask)])` step on its own (text.rs:153-162)
  std::vector<float> embed_ids(const std::vector<int64_t>& ids, const std::vector<int64_t>* mask, size_t batch) const {
    const size_t ctx = context_length();
    if (batch == 0) throw ClipError::Inference("Empty batch");
    if (ids.size() != batch * ctx || (mask != nullptr && mask->size() != ids.size()))
      throw ClipError::Shape("input_ids must be [batch, context_length]");
    std::vector<float> out(batch * static_cast<size_t>(session.embed_dim()));
    std::lock_guard<std::mutex> guard(session.write_lock());
    OnnxSession::check(clipb200_text_embed(session.handle(), ids.data(), mask ? mask->data() : nullptr,
                                           static_cast<int64_t>(batch), static_cast<int64_t>(ctx), out.data()));
    return out;
  }

 private:
  TextEmbedder(OnnxSession s, OpenClipConfig c, ModelConfig m, std::string dir, std::string id)
      : session(std::move(s)), config(std::move(c)), model_config(std::move(m)), model_dir(std::move(dir)), id_name(std::move(id)) {}
};

// ------------------------------------------------------------------------------------------------ clip
class Clip {  // src/clip.rs
 public:
  VisionEmbedder vision;
  TextEmbedder text;
  std::string model_dir;

  static Clip from_local_dir(const std::string& model_dir, int cuda_device = 0) {  // clip.rs:51-66
    model_manager::verify_model_dir(model_dir);
    return Clip(VisionEmbedder::from_local_dir(model_dir, cuda_device), TextEmbedder::from_local_dir(model_dir, cuda_device), model_dir);
  }
  static Clip from_local_id(const std::string& model_id, const std::string& base_folder = "", int cuda_device = 0) {  // clip.rs:37-48
    return from_local_dir(detail::join(base_folder.empty() ? model_manager::get_default_base_folder() : base_folder, model_id), cuda_device);
  }
  Clip duplicate() const {  // clip.rs:69-73
    Clip c = from_local_dir(model_dir, vision.session.device());
    c.text.encoder = text.encoder;
    return c;
  }
  ModelConfig get_model_config() const { return text.model_config; }  // clip.rs:75-77

  float compare(const RgbImage& image, const std::string& txt) const {  // clip.rs:81-90
    const std::vector<float> v = vision.embed_image(image), t = text.embed_text(txt);
    return fmaf(dot(v.data(), t.data(), v.size()), scale(), bias());
  }
  std::vector<std::pair<std::string, float>> classify(const RgbImage& image, const std::vector<std::string>& labels) const {  // clip.rs:94-132
    const std::vector<float> v = vision.embed_image(image), t = text.embed_texts(labels);
    const std::vector<float> probs = probabilities(t, v, labels.size());
    std::vector<std::pair<std::string, float>> results;
    for (size_t i = 0; i < labels.size(); ++i) results.emplace_back(labels[i], probs[i]);
    std::stable_sort(results.begin(), results.end(), [](const auto& a, const auto& b) { return a.second > b.second; });
    return results;
  }
  std::vector<std::pair<size_t, float>> rank_images(const std::vector<RgbImage>& images, const std::string& txt) const {  // clip.rs:136-170
    const std::vector<float> embs = vision.embed_images(images), t = text.embed_text(txt);
    const std::vector<float> probs = probabilities(embs, t, images.size());
    std::vector<std::pair<size_t, float>> results;
    for (size_t i = 0; i < probs.size(); ++i) results.emplace_back(i, probs[i]);
    std::stable_sort(results.begin(), results.end(), [](const auto& a, const auto& b) { return a.second > b.second; });
    return results;
  }
  static std::vector<float> softmax(const std::vector<float>& logits) {  // clip.rs:174-179
    float mx = -INFINITY;
    for (float x : logits) mx = fmaxf(mx, x);
    std::vector<float> e(logits.size());
    float sum = 0.f;
    for (size_t i = 0; i < logits.size(); ++i) { e[i] = expf(logits[i] - mx); sum += e[i]; }
    for (float& x : e) x /= sum;
    return e;
  }
  static float sigmoid(float logit) { return 1.0f / (1.0f + expf(-logit)); }  // clip.rs:183-185

 private:
  Clip(VisionEmbedder v, TextEmbedder t, std::string dir) : vision(std::move(v)), text(std::move(t)), model_dir(std::move(dir)) {}
  float scale() const { return text.model_config.logit_scale.value_or(1.0f); }
  float bias() const { return text.model_config.logit_bias.value_or(0.0f); }
  static float dot(const float* a, const float* b, size_t n) {
    float s = 0.f;
    for (size_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
  }
  // rows [n, D] . query [D] -> mul_add(scale, bias) -> sigmoid per row | softmax over rows (clip.rs:102-121)
  std::vector<float> probabilities(const std::vector<float>& rows, const std::vector<float>& query, size_t n) const {
    const size_t D = query.size();
    std::vector<float> logits(n);
    for (size_t i = 0; i < n; ++i) logits[i] = fmaf(dot(rows.data() + i * D, query.data(), D), scale(), bias());
    if (text.model_config.activation_function.value_or("softmax") == "sigmoid") {
      for (float& l : logits) l = sigmoid(l);
      return logits;
    }
    return softmax(logits);
  }
};

}  // namespace clipb200

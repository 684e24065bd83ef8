// Exercises include/clipb200.hpp (the C++ host mirror of the open_clip_inference API).
//   host_api_test errors <scratch_dir>                 no GPU: error kinds / messages, config parsing, softmax / sigmoid
//   host_api_test run <model_dir> <images.u8> <n> <w> <h> <ids.i64> <lens.i64> <n_texts>
//                                                      GPU: prints embeddings / classify / rank / compare as JSON
#include <stdio.h>
#include <string.h>

#include <fstream>
#include <iostream>

#include "clipb200.hpp"

using namespace clipb200;

static void write_file(const std::string& path, const std::string& body) {
  std::ofstream f(path, std::ios::binary);
  f << body;
}

template <class F>
static bool throws(ErrorKind kind, const std::string& needle, F f) {
  try {
    f();
  } catch (const ClipError& e) {
    if (e.kind() == kind && std::string(e.what()).find(needle) != std::string::npos) return true;
    fprintf(stderr, "wrong error: kind %d '%s' (wanted '%s')\n", static_cast<int>(e.kind()), e.what(), needle.c_str());
    return false;
  }
  fprintf(stderr, "no error raised (wanted '%s')\n", needle.c_str());
  return false;
}

static int run_errors(const std::string& dir) {
  int bad = 0;
  bad += !throws(ErrorKind::ModelFolderNotFound, "Model folder not found, generate it with `uv run pull_onnx.py -h`. '" + dir + "/nope'",
                 [&] { model_manager::verify_model_dir(dir + "/nope"); });
  bad += !throws(ErrorKind::MissingModelFile, "Missing model file 'model_config.json' in folder '" + dir + "'",
                 [&] { model_manager::verify_model_dir(dir); });
  bad += !throws(ErrorKind::ModelFolderNotFound, "nope", [&] { Clip::from_local_dir(dir + "/nope"); });
  write_file(dir + "/model_config.json",
             "{\"logit_scale\": 112.5, \"logit_bias\": -16.25, \"activation_function\": \"sigmoid\", "
             "\"tokenizer_needs_lowercase\": true, \"pad_id\": 1, \"vocab_size\": 256000}");
  const ModelConfig mc = ModelConfig::from_file(dir + "/model_config.json");
  bad += !(mc.tokenizer_needs_lowercase && mc.activation_function.value() == "sigmoid" && mc.logit_scale.value() == 112.5f &&
           mc.logit_bias.value() == -16.25f && mc.pad_id.value() == 1u);
  write_file(dir + "/empty_model_config.json", "{}");
  const ModelConfig me = ModelConfig::from_file(dir + "/empty_model_config.json");  // all fields optional / defaulted
  bad += !(!me.tokenizer_needs_lowercase && !me.activation_function && !me.logit_scale && !me.logit_bias && !me.pad_id);
  write_file(dir + "/open_clip_config.json",
             "{\"model_cfg\": {\"embed_dim\": 512, \"vision_cfg\": {\"image_size\": 224, \"layers\": 12, \"width\": 768, "
             "\"patch_size\": 32}, \"text_cfg\": {\"context_length\": 77, \"vocab_size\": 49408}, \"quick_gelu\": true},\n"
             " \"preprocess_cfg\": {\"mean\": [0.48145466, 0.4578275, 0.40821073], \"std\": [0.26862954, 0.26130258, 0.27577711]}}");
  const OpenClipConfig oc = OpenClipConfig::from_file(dir + "/open_clip_config.json");
  bad += !(oc.model_cfg.embed_dim == 512 && oc.model_cfg.vision_cfg.image_size == 224 && oc.model_cfg.vision_cfg.layers.value() == 12 &&
           oc.model_cfg.text_cfg.context_length == 77 && !oc.model_cfg.text_cfg.hf_tokenizer_name &&
           oc.preprocess_cfg.interpolation == "bicubic" && oc.preprocess_cfg.resize_mode == "shortest" &&
           oc.preprocess_cfg.mean[0] == 0.48145466f && oc.preprocess_cfg.std[2] == 0.27577711f);
  write_file(dir + "/bad.json", "{\"model_cfg\": {\"embed_dim\": 512}}");
  bad += !throws(ErrorKind::Json, "missing field `vision_cfg`", [&] { OpenClipConfig::from_file(dir + "/bad.json"); });
  write_file(dir + "/worse.json", "{\"model_cfg\": ");
  bad += !throws(ErrorKind::Json, "EOF", [&] { OpenClipConfig::from_file(dir + "/worse.json"); });
  bad += !throws(ErrorKind::Io, "nothing.json", [&] { ModelConfig::from_file(dir + "/nothing.json"); });
  const std::vector<float> p = Clip::softmax({1.0f, 2.0f, 3.0f});
  bad += !(fabsf(p[0] - 0.09003057f) < 1e-6f && fabsf(p[2] - 0.66524096f) < 1e-6f && fabsf(Clip::sigmoid(0.5f) - 0.62245935f) < 1e-6f);
  bad += !(model_manager::get_default_base_folder().find("open_clip") != std::string::npos);
  printf("%s\n", bad ? "HOST API ERRORS FAILED" : "HOST API ERRORS PASSED");
  return bad;
}

template <class T>
static std::vector<T> read_bin(const std::string& path, size_t count) {
  std::vector<T> v(count);
  std::ifstream f(path, std::ios::binary);
  f.read(reinterpret_cast<char*>(v.data()), static_cast<std::streamsize>(count * sizeof(T)));
  if (static_cast<size_t>(f.gcount()) != count * sizeof(T)) throw ClipError::Io("short read: " + path);
  return v;
}

static void print_floats(const char* key, const std::vector<float>& v, bool last = false) {
  printf("\"%s\": [", key);
  for (size_t i = 0; i < v.size(); ++i) printf("%s%.9g", i ? ", " : "", v[i]);
  printf("]%s\n", last ? "" : ",");
}

static int run_model(int argc, char** argv) {
  if (argc < 10) return 2;
  const std::string model_dir = argv[2];
  const size_t n = strtoul(argv[4], nullptr, 10);
  const int w = atoi(argv[5]), h = atoi(argv[6]);
  const size_t nt = strtoul(argv[9], nullptr, 10);
  Clip clip = Clip::from_local_dir(model_dir);
  const size_t ctx = clip.text.context_length();
  const std::vector<uint8_t> pixels = read_bin<uint8_t>(argv[3], n * w * h * 3);
  const std::vector<int64_t> ids = read_bin<int64_t>(argv[7], nt * ctx), lens = read_bin<int64_t>(argv[8], nt);
  std::vector<RgbImage> images;
  for (size_t i = 0; i < n; ++i) images.push_back({pixels.data() + i * static_cast<size_t>(w) * h * 3, w, h});
  std::vector<std::string> labels;
  for (size_t i = 0; i < nt; ++i) labels.push_back("#" + std::to_string(i));
  // stand-in for the `tokenizers` crate: label "#k" -> row k of the id file (ids with special tokens, unpadded)
  clip.text.encoder = [&](const std::string& s) {
    const size_t k = strtoul(s.c_str() + 1, nullptr, 10);
    std::vector<uint32_t> out;
    for (int64_t j = 0; j < lens[k]; ++j) out.push_back(static_cast<uint32_t>(ids[k * ctx + j]));
    return out;
  };
  printf("{\"embed_dim\": %lld, \"input_name\": \"%s\", \"id_name\": \"%s\", \"has_mask\": %s,\n",
         static_cast<long long>(clip.vision.session.embed_dim()), clip.vision.input_name.c_str(), clip.text.id_name.c_str(),
         clip.text.mask_name ? "true" : "false");
  print_floats("image_embeddings", clip.vision.embed_images(images));
  print_floats("text_embeddings", clip.text.embed_texts(labels));
  print_floats("preprocess0", [&] { auto p = clip.vision.preprocess(images[0]); p.resize(64); return p; }());
  const auto cls = clip.classify(images[0], labels);
  printf("\"classify\": [");
  for (size_t i = 0; i < cls.size(); ++i) printf("%s[\"%s\", %.9g]", i ? ", " : "", cls[i].first.c_str(), cls[i].second);
  printf("],\n");
  const auto rank = clip.rank_images(images, labels[0]);
  printf("\"rank_images\": [");
  for (size_t i = 0; i < rank.size(); ++i) printf("%s[%zu, %.9g]", i ? ", " : "", rank[i].first, rank[i].second);
  printf("],\n");
  bool empty_ok = false;
  try { clip.vision.embed_images({}); } catch (const ClipError& e) { empty_ok = e.kind() == ErrorKind::Inference && std::string(e.what()) == "Inference error: Empty batch"; }
  Clip twin = clip.duplicate();  // clip.rs:69-73: an independent session on the same files
  const float c0 = clip.compare(images[0], labels[1]), c1 = twin.compare(images[0], labels[1]);
  printf("\"empty_batch_error\": %s, \"duplicate_matches\": %s, \"compare\": %.9g}\n", empty_ok ? "true" : "false",
         c0 == c1 ? "true" : "false", c0);
  return 0;
}

int main(int argc, char** argv) {
  try {
    if (argc >= 3 && strcmp(argv[1], "errors") == 0) return run_errors(argv[2]);
    if (argc >= 2 && strcmp(argv[1], "run") == 0) return run_model(argc, argv);
  } catch (const ClipError& e) {
    fprintf(stderr, "ClipError(kind %d): %s\n", static_cast<int>(e.kind()), e.what());
    return 3;
  }
  fprintf(stderr, "usage: host_api_test errors <dir> | run <model_dir> <images.u8> <n> <w> <h> <ids.i64> <lens.i64> <n_texts>\n");
  return 2;
}

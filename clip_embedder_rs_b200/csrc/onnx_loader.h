// Minimal ONNX reader: protobuf wire format -> initializer table + graph input/output names + metadata_props.
// Replaces what `Session::builder().commit_from_file(path)` does for the reference (src/onnx.rs:19-23) as far as
// this engine needs it: it never builds an executable graph, it only binds initializers to kernel operands.
// External data (`<file>.onnx.data`, data_location = EXTERNAL) is memory-mapped relative to the .onnx file's
// directory, which is how the reference's model directories store weights (src/model_manager.rs:16-17).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

namespace clipb200 {

struct MappedFile {
  void* base = nullptr;
  size_t size = 0;
  ~MappedFile();
};

struct OnnxTensor {
  std::string name;
  std::vector<int64_t> dims;
  int data_type = 0;  // 1 f32, 7 i64, 10 f16, 16 bf16
  const uint8_t* data = nullptr;  // points into a mapped file, or into `owned`
  size_t nbytes = 0;
  std::vector<uint8_t> owned;  // float_data / int64_data fields converted to raw little-endian
  // Set by the graph recogniser (onnx_graph.cc) on canonical-name aliases of 2-D weights: `dims` are the canonical
  // [rows, cols] but the bytes are stored as the transpose ([cols, rows] row-major), e.g. a `Linear` weight that the
  // exporter pre-transposed into a MatMul operand.
  bool transposed = false;
  int64_t numel() const {  // -1 for malformed dims (negative / overflowing): every size comparison then fails closed
    int64_t n = 1;
    for (int64_t d : dims) {
      if (d < 0) return -1;
      if (d > 0 && n > (int64_t(1) << 50) / d) return -1;
      n *= d;
    }
    return n;
  }
};

struct OnnxAttr {
  int type = 0;  // AttributeProto.type: 1 FLOAT, 2 INT, 3 STRING, 4 TENSOR, 6 FLOATS, 7 INTS
  float f = 0.f;
  int64_t i = 0;
  std::string s;
  std::vector<int64_t> ints;
  std::vector<float> floats;
  std::shared_ptr<OnnxTensor> t;
};

struct OnnxNode {
  std::string op_type, name;
  std::vector<std::string> inputs, outputs;
  std::map<std::string, OnnxAttr> attrs;
  const OnnxAttr* attr(const std::string& k) const {
    auto it = attrs.find(k);
    return it == attrs.end() ? nullptr : &it->second;
  }
  int64_t attr_i(const std::string& k, int64_t dflt) const {
    const OnnxAttr* a = attr(k);
    return a ? a->i : dflt;
  }
  float attr_f(const std::string& k, float dflt) const {
    const OnnxAttr* a = attr(k);
    return a ? a->f : dflt;
  }
};

struct OnnxValueInfo {
  std::string name;
  int elem_type = 0;
  std::vector<int64_t> dims;  // -1 for symbolic (dim_param) dimensions
};

struct OnnxModel {
  std::vector<std::string> inputs;   // graph inputs that are not initializers
  std::vector<std::string> outputs;
  std::vector<OnnxValueInfo> input_infos;  // same order as `inputs`
  std::map<std::string, OnnxTensor> initializers;
  std::map<std::string, std::string> metadata;
  std::vector<OnnxNode> nodes;
  int64_t opset = 0;
  std::vector<std::shared_ptr<MappedFile>> files;  // keeps mappings alive

  const OnnxTensor* find(const std::string& name) const;
  bool has(const std::string& name) const { return find(name) != nullptr; }
  std::string meta(const std::string& key, const std::string& dflt = "") const;
};

// Returns false and fills `err` on failure.
bool load_onnx(const std::string& path, OnnxModel* model, std::string* err);

}  // namespace clipb200

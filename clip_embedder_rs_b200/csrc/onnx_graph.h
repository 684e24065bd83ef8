// Name-independent binding of exported CLIP / SigLIP towers.
//
// The reference hands `visual.onnx` / `text.onnx` to onnxruntime, which executes whatever graph `torch.onnx.export`
// wrote (`/root/reference/pull_onnx.py:169-181`, `/root/reference/src/onnx.rs:19-23`).  Such graphs do not keep the
// open_clip / timm parameter names for everything the engine needs: `Linear` weights become pre-transposed
// `onnx::MatMul_<n>` initializers, identical tensors are de-duplicated, the head count only exists inside Reshape
// arithmetic, an optimiser may rename every initializer.  This module recovers the architecture from the graph itself:
//
//   1. an abstract interpreter walks the nodes with batch = 1, giving every tensor a shape, a "depends on the graph
//      input" bit, and — for input-independent tensors up to 8 Mi elements — its value (constant folding: shape
//      arithmetic, causal masks, the class token, the attention-pool query);
//   2. the input-dependent nodes are reduced to a token stream (patch conv | token gather, class-token concat,
//      positional add, LayerNorm, linear sites with their bias, softmax sites with heads / scale / mask / q,k,v
//      provenance, activation kind, token select, L2 norm);
//   3. a small grammar matches the stream against the tower layouts the engine implements (pre-norm ViT / text
//      transformer with fused or split q,k,v; class-token + projection head or attention-pool (MAP) head; argmax /
//      last / first token pooling) and emits the tensors under the canonical names `Engine::LoadVision/LoadText` bind,
//      plus the `clipb200.*` hyper-parameters (heads, activation, eps, pooling, causal) as metadata.
//
// Nothing here touches the GPU; `clipb200_onnx_inspect` exposes the result for CPU-only tests.
#pragma once
#include <string>
#include <vector>

#include "onnx_loader.h"

namespace clipb200 {

struct GraphBinding {
  std::string canonical;  // name the engine binds
  std::string source;     // initializer it came from, or a description ("<folded constant>", "concat(a,b,c)")
  bool transposed = false;
};

// True when the file carries an executable graph (MatMul/Gemm/Conv nodes) and was not written by
// tools/export_synthetic.py (which states its hyper-parameters in `clipb200.*` metadata).
bool graph_needs_recognition(const OnnxModel& m);

// On success adds canonical-named initializers and `clipb200.*` metadata to `m` and returns true.
// On failure returns false with a reason in `err`; `m` is left untouched.
bool recognize_graph(OnnxModel* m, std::string* err, std::vector<GraphBinding>* bindings_or_null);

// Re-parameterised FastViT trunk exported as a real graph: registers the attention blocks' two Linear weights (renamed
// and pre-transposed by the exporter) under timm's names by following the graph from the named tensors beside them.
// No-op for initializer-only files.  Returns false with a reason when an attention block cannot be resolved.
bool bind_fastvit_graph(OnnxModel* m, std::string* err);

// fp32 copy of an initializer in its canonical layout (undoes `OnnxTensor::transposed`); f32 / f16 / bf16 sources.
bool tensor_to_f32(const OnnxTensor& t, std::vector<float>* out);

}  // namespace clipb200

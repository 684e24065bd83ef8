// Parses + structurally recognises every file named on the command line (no GPU).  Built with
// -fsanitize=address,undefined by tests/test_onnx_fuzz_cpu.py and fed byte-mutated real exports: the loader and the
// graph recogniser sit behind a C ABI that takes user-supplied files, so malformed input must be declined, not crash.
#include <stdio.h>
#include "onnx_graph.h"
int main(int argc, char** argv) {
  int loaded = 0, recognised = 0, fastvit = 0;
  for (int i = 1; i < argc; ++i) {
    clipb200::OnnxModel m;
    std::string err;
    if (!clipb200::load_onnx(argv[i], &m, &err)) continue;
    ++loaded;
    std::vector<clipb200::GraphBinding> b;
    if (clipb200::graph_needs_recognition(m) && clipb200::recognize_graph(&m, &err, &b)) {
      ++recognised;
      std::vector<float> v;  // touch every bound tensor the way the engine would
      for (const clipb200::GraphBinding& g : b)
        if (const clipb200::OnnxTensor* t = m.find(g.canonical)) clipb200::tensor_to_f32(*t, &v);
    } else if (m.has("model.visual.trunk.stem.0.reparam_conv.weight")) {
      // the FastViT route of Engine::Init: attention Linears located through the graph's edges
      std::string fv_err;
      if (clipb200::bind_fastvit_graph(&m, &fv_err)) {
        ++fastvit;
        std::vector<float> v;
        for (const auto& kv : m.initializers)
          if (kv.second.transposed) clipb200::tensor_to_f32(kv.second, &v);
      }
    }
  }
  printf("FUZZ HARNESS DONE files=%d loaded=%d recognised=%d fastvit=%d\n", argc - 1, loaded, recognised, fastvit);
  return 0;
}

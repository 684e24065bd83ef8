#include "onnx_graph.h"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <set>
#include <unordered_map>

namespace clipb200 {
namespace {

constexpr int64_t kMaxConst = int64_t(1) << 23;  // largest input-independent tensor that is folded to a value
using Shape = std::vector<int64_t>;

constexpr int64_t kNumelCap = int64_t(1) << 52;  // saturation value: far above anything that is ever materialised

// Saturating product; a negative dimension (malformed file) also saturates so that every size check fails closed.
int64_t numel(const Shape& s) {
  int64_t n = 1;
  for (int64_t d : s) {
    if (d < 0) return kNumelCap;
    if (d == 0) return 0;
    if (n > kNumelCap / d) return kNumelCap;
    n *= d;
  }
  return n;
}

bool shape_ok(const Shape& s) {
  if (s.size() > 8) return false;
  for (int64_t d : s) if (d < 0 || d > (int64_t(1) << 40)) return false;
  return numel(s) < kNumelCap;
}

std::string shape_str(const Shape& s) {
  std::string o = "[";
  for (size_t i = 0; i < s.size(); ++i) o += (i ? "," : "") + std::to_string(s[i]);
  return o + "]";
}

bool is_float_dt(int dt) { return dt == 1 || dt == 10 || dt == 11 || dt == 16; }

float half_bits_to_float(uint16_t h) {
  const uint32_t sign = (h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1F, man = h & 0x3FF, bits;
  if (exp == 0) {
    if (man == 0) bits = sign;
    else {
      exp = 127 - 15 + 1;
      while (!(man & 0x400)) { man <<= 1; --exp; }
      bits = sign | (exp << 23) | ((man & 0x3FF) << 13);
    }
  } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
  else bits = sign | ((exp - 15 + 127) << 23) | (man << 13);
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

bool read_tensor(const OnnxTensor& t, std::vector<double>* out) {
  const int64_t n = t.numel();
  if (n < 0 || n > (int64_t(1) << 31)) return false;
  out->resize(static_cast<size_t>(n));
  if (n == 0) return true;
  if (t.data == nullptr) return false;
  const uint8_t* p = t.data;
  switch (t.data_type) {
    case 1: { if (t.nbytes < static_cast<size_t>(n) * 4) return false; for (int64_t i = 0; i < n; ++i) { float v; memcpy(&v, p + 4 * i, 4); (*out)[i] = v; } return true; }
    case 11: { if (t.nbytes < static_cast<size_t>(n) * 8) return false; for (int64_t i = 0; i < n; ++i) { double v; memcpy(&v, p + 8 * i, 8); (*out)[i] = v; } return true; }
    case 7: { if (t.nbytes < static_cast<size_t>(n) * 8) return false; for (int64_t i = 0; i < n; ++i) { int64_t v; memcpy(&v, p + 8 * i, 8); (*out)[i] = static_cast<double>(v); } return true; }
    case 6: { if (t.nbytes < static_cast<size_t>(n) * 4) return false; for (int64_t i = 0; i < n; ++i) { int32_t v; memcpy(&v, p + 4 * i, 4); (*out)[i] = v; } return true; }
    case 9: case 2: { if (t.nbytes < static_cast<size_t>(n)) return false; for (int64_t i = 0; i < n; ++i) (*out)[i] = p[i]; return true; }
    case 3: { if (t.nbytes < static_cast<size_t>(n)) return false; for (int64_t i = 0; i < n; ++i) (*out)[i] = static_cast<int8_t>(p[i]); return true; }
    case 10: { if (t.nbytes < static_cast<size_t>(n) * 2) return false; for (int64_t i = 0; i < n; ++i) { uint16_t h; memcpy(&h, p + 2 * i, 2); (*out)[i] = half_bits_to_float(h); } return true; }
    case 16: { if (t.nbytes < static_cast<size_t>(n) * 2) return false; for (int64_t i = 0; i < n; ++i) { uint16_t h; memcpy(&h, p + 2 * i, 2); const uint32_t b = static_cast<uint32_t>(h) << 16; float f; memcpy(&f, &b, 4); (*out)[i] = f; } return true; }
    default: return false;
  }
}

// Abstract value of one graph tensor.
struct AVal {
  Shape shape;
  int dtype = 1;
  bool tainted = false;   // value depends on a graph input (shapes never do: batch is fixed to 1)
  bool has_data = false;  // `data` holds the value (input-independent tensors only)
  std::vector<double> data;
  const OnnxTensor* init = nullptr;  // the value is exactly this initializer / Constant (same element order) ...
  bool init_t = false;               // ... or its 2-D transpose
  std::string init_name;
  int producer = -1;
};

bool materialize(AVal* v) {
  if (v->has_data) return true;
  if (v->tainted || v->init == nullptr) return false;
  const int64_t n = numel(v->shape);
  if (n > kMaxConst || n != v->init->numel()) return false;
  if (!read_tensor(*v->init, &v->data)) return false;
  if (v->init_t) {
    const int64_t r = v->init->dims[0], c = v->init->dims[1];
    std::vector<double> t(v->data.size());
    for (int64_t i = 0; i < r; ++i)
      for (int64_t j = 0; j < c; ++j) t[static_cast<size_t>(j * r + i)] = v->data[static_cast<size_t>(i * c + j)];
    v->data.swap(t);
  }
  v->has_data = true;
  return true;
}

int64_t to_i64(double d) {
  if (d > 4.0e18) return INT64_MAX / 2;
  if (d < -4.0e18) return INT64_MIN / 2;
  return static_cast<int64_t>(d);
}

Shape strides_of(const Shape& s) {
  Shape st(s.size(), 1);
  for (int i = static_cast<int>(s.size()) - 2; i >= 0; --i) st[i] = st[i + 1] * s[i + 1];
  return st;
}

bool bshape(const Shape& a, const Shape& b, Shape* o) {
  const size_t r = std::max(a.size(), b.size());
  o->assign(r, 1);
  for (size_t i = 0; i < r; ++i) {
    const int64_t da = i < r - a.size() ? 1 : a[i - (r - a.size())];
    const int64_t db = i < r - b.size() ? 1 : b[i - (r - b.size())];
    if (da != db && da != 1 && db != 1) return false;
    (*o)[i] = da == 1 ? db : da;
  }
  return true;
}

// strides of `in` aligned to the rank of `out`, 0 along broadcast dimensions
Shape bstrides(const Shape& in, const Shape& out) {
  Shape st(out.size(), 0);
  const Shape s = strides_of(in);
  const size_t off = out.size() - in.size();
  for (size_t i = 0; i < in.size(); ++i) st[i + off] = in[i] == 1 ? 0 : s[i];
  return st;
}

// calls f(linear_out_index, multi_index) for every element of `shape`
template <class F>
void for_each(const Shape& shape, F f) {
  const int64_t n = numel(shape);
  Shape idx(shape.size(), 0);
  for (int64_t i = 0; i < n; ++i) {
    f(i, idx);
    for (int d = static_cast<int>(shape.size()) - 1; d >= 0; --d) {
      if (++idx[d] < shape[d]) break;
      idx[d] = 0;
    }
  }
}

int64_t dot(const Shape& idx, const Shape& st) {
  int64_t o = 0;
  for (size_t i = 0; i < idx.size(); ++i) o += idx[i] * st[i];
  return o;
}

// ----------------------------------------------------------------------------------------------------------
// 1. abstract interpretation
// ----------------------------------------------------------------------------------------------------------
struct Analyzer {
  const OnnxModel& m;
  std::unordered_map<std::string, AVal> env;
  std::string err;
  int64_t batch0 = 1;  // batch the graph is analysed at: 1 for a dynamic batch axis, else the exported batch

  explicit Analyzer(const OnnxModel& model) : m(model) {}

  bool fail(const std::string& s) {
    if (err.empty()) err = s;
    return false;
  }
  AVal* get(const std::string& name) {
    if (name.empty()) return nullptr;
    auto it = env.find(name);
    return it == env.end() ? nullptr : &it->second;
  }
  bool ints_of(AVal* v, std::vector<int64_t>* out) {
    if (v == nullptr || !materialize(v)) return false;
    out->clear();
    for (double d : v->data) out->push_back(to_i64(d));
    return true;
  }

  bool init_env() {
    for (const auto& kv : m.initializers) {
      AVal v;
      v.shape = kv.second.dims;
      v.dtype = kv.second.data_type;
      v.init = &kv.second;
      v.init_name = kv.first;
      env.emplace(kv.first, std::move(v));
    }
    for (const OnnxValueInfo& vi : m.input_infos) {
      AVal v;
      v.shape = vi.dims;
      for (size_t i = 0; i < v.shape.size(); ++i)
        if (v.shape[i] <= 0) {
          if (i != 0) return fail("graph input '" + vi.name + "' has a dynamic dimension other than the batch axis");
          v.shape[i] = 1;
        }
      // a fixed-batch export (no dynamic_axes) bakes its batch into Reshape constants: analyse it at that batch
      if (!v.shape.empty()) batch0 = v.shape[0];
      v.dtype = vi.elem_type ? vi.elem_type : 1;
      v.tainted = true;
      env[vi.name] = std::move(v);
    }
    return true;
  }

  bool run() {
    if (!init_env()) return false;
    for (const auto& kv : env)
      if (!shape_ok(kv.second.shape)) return fail("tensor '" + kv.first + "' has an invalid shape " + shape_str(kv.second.shape));
    for (size_t i = 0; i < m.nodes.size(); ++i) {
      bool ok = eval(static_cast<int>(i));
      if (ok)
        for (const std::string& o : m.nodes[i].outputs) {
          AVal* v = get(o);
          if (v != nullptr && !shape_ok(v->shape)) ok = fail("output '" + o + "' has an invalid shape " + shape_str(v->shape));
        }
      if (!ok) {
        if (err.empty()) err = "graph analysis failed";
        err += " (node " + std::to_string(i) + " " + m.nodes[i].op_type + " '" + m.nodes[i].name + "')";
        return false;
      }
    }
    return true;
  }

  void put(const OnnxNode& n, int idx, AVal v, size_t out = 0) {
    if (out >= n.outputs.size() || n.outputs[out].empty()) return;
    v.producer = idx;
    env[n.outputs[out]] = std::move(v);
  }

  bool axes_arg(const OnnxNode& n, size_t input_idx, std::vector<int64_t>* axes, bool* present) {
    *present = false;
    if (n.inputs.size() > input_idx && !n.inputs[input_idx].empty()) {
      if (!ints_of(get(n.inputs[input_idx]), axes)) return fail("axes input is not a constant");
      *present = true;
    } else if (const OnnxAttr* a = n.attr("axes")) {
      *axes = a->ints;
      *present = true;
    }
    return true;
  }

  bool eval(int idx) {
    const OnnxNode& n = m.nodes[idx];
    const std::string& op = n.op_type;
    std::vector<AVal*> in;
    for (const std::string& s : n.inputs) {
      AVal* v = get(s);
      if (!s.empty() && v == nullptr) return fail("input '" + s + "' is undefined");
      in.push_back(v);
    }
    auto need = [&](size_t k) { return in.size() > k && in[k] != nullptr; };
    bool any_taint = false;
    for (AVal* v : in) any_taint |= (v != nullptr && v->tainted);

    if (op == "Constant") {
      AVal v;
      if (const OnnxAttr* a = n.attr("value")) {
        if (!a->t) return fail("Constant without an inline tensor");
        v.shape = a->t->dims;
        v.dtype = a->t->data_type;
        v.init = a->t.get();
      } else if (const OnnxAttr* a = n.attr("value_float")) {
        v.dtype = 1; v.has_data = true; v.data = {a->f};
      } else if (const OnnxAttr* a = n.attr("value_int")) {
        v.dtype = 7; v.has_data = true; v.data = {static_cast<double>(a->i)};
      } else if (const OnnxAttr* a = n.attr("value_ints")) {
        v.dtype = 7; v.has_data = true; v.shape = {static_cast<int64_t>(a->ints.size())};
        for (int64_t x : a->ints) v.data.push_back(static_cast<double>(x));
      } else if (const OnnxAttr* a = n.attr("value_floats")) {
        v.dtype = 1; v.has_data = true; v.shape = {static_cast<int64_t>(a->floats.size())};
        for (float x : a->floats) v.data.push_back(x);
      } else return fail("Constant with an unsupported value attribute");
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "Identity" || op == "Dropout") {
      if (!need(0)) return fail("missing input");
      AVal v = *in[0];
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "Shape") {
      if (!need(0)) return fail("missing input");
      const int64_t r = static_cast<int64_t>(in[0]->shape.size());
      int64_t s = n.attr_i("start", 0), e = n.attr_i("end", r);
      if (s < 0) s += r;
      if (e < 0) e += r;
      s = std::max<int64_t>(0, std::min(s, r));
      e = std::max<int64_t>(0, std::min(e, r));
      AVal v;
      v.dtype = 7; v.has_data = true;
      for (int64_t i = s; i < e; ++i) v.data.push_back(static_cast<double>(in[0]->shape[i]));
      v.shape = {static_cast<int64_t>(v.data.size())};
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "Cast") {
      if (!need(0)) return fail("missing input");
      AVal v = *in[0];
      const int to = static_cast<int>(n.attr_i("to", 1));
      const bool to_float = is_float_dt(to);
      if (!v.tainted) {
        const bool keep_alias = to_float && is_float_dt(in[0]->dtype) && !v.has_data;  // f16 -> f32 of an initializer
        if (!keep_alias) {
          if (materialize(&v)) {
            for (double& d : v.data) {
              if (to == 9) d = d != 0.0;
              else if (!to_float) d = static_cast<double>(to_i64(d));
              else if (to == 1) d = static_cast<double>(static_cast<float>(d));
            }
          }
          v.init = nullptr;
          v.init_t = false;
        }
      }
      v.dtype = to;
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "Reshape" || op == "Flatten" || op == "Unsqueeze" || op == "Squeeze") {
      if (!need(0)) return fail("missing input");
      AVal v = *in[0];
      const Shape& s = in[0]->shape;
      Shape o;
      if (op == "Reshape") {
        std::vector<int64_t> t;
        if (!need(1) || !ints_of(in[1], &t)) return fail("Reshape target is not a constant");
        const bool allowzero = n.attr_i("allowzero", 0) != 0;
        int64_t known = 1;
        int neg = -1;
        for (size_t i = 0; i < t.size(); ++i) {
          int64_t d = t[i];
          if (d == 0 && !allowzero) {
            if (i >= s.size()) return fail("Reshape: 0 beyond input rank");
            d = s[i];
          }
          if (d == -1) {
            if (neg >= 0) return fail("Reshape: more than one -1");
            neg = static_cast<int>(i);
          } else {
            if (d < 0 || d > (int64_t(1) << 40) || (d > 0 && known > kNumelCap / d)) return fail("Reshape: invalid target dimension");
            known *= d;
          }
          o.push_back(d);
        }
        if (neg >= 0) {
          if (known == 0 || numel(s) % known != 0) return fail("Reshape: cannot infer -1");
          o[neg] = numel(s) / known;
        }
        if (numel(o) != numel(s)) return fail("Reshape " + shape_str(s) + " -> " + shape_str(o) + ": element count differs");
      } else if (op == "Flatten") {
        int64_t ax = n.attr_i("axis", 1);
        if (ax < 0) ax += static_cast<int64_t>(s.size());
        if (ax < 0 || ax > static_cast<int64_t>(s.size())) return fail("Flatten: axis out of range");
        int64_t a = 1, b = 1;
        for (size_t i = 0; i < s.size(); ++i) (static_cast<int64_t>(i) < ax ? a : b) *= s[i];
        o = {a, b};
      } else {
        std::vector<int64_t> axes;
        bool present;
        if (!axes_arg(n, 1, &axes, &present)) return false;
        if (op == "Unsqueeze") {
          if (!present) return fail("Unsqueeze without axes");
          const int64_t r = static_cast<int64_t>(s.size() + axes.size());
          std::set<int64_t> ax;
          for (int64_t a : axes) {
            const int64_t p = a < 0 ? a + r : a;
            if (p < 0 || p >= r || !ax.insert(p).second) return fail("Unsqueeze: axis out of range or repeated");
          }
          size_t k = 0;
          for (int64_t i = 0; i < r; ++i) o.push_back(ax.count(i) ? 1 : s[k++]);
        } else {
          std::set<int64_t> ax;
          for (int64_t a : axes) {
            const int64_t p = a < 0 ? a + static_cast<int64_t>(s.size()) : a;
            if (p < 0 || p >= static_cast<int64_t>(s.size())) return fail("Squeeze: axis out of range");
            ax.insert(p);
          }
          for (size_t i = 0; i < s.size(); ++i) {
            const bool drop = present ? ax.count(static_cast<int64_t>(i)) > 0 : s[i] == 1;
            if (!drop) o.push_back(s[i]);
          }
        }
      }
      v.shape = o;
      if (v.init_t) {  // a reshaped transpose is no longer a plain alias
        if (!materialize(&v)) { v.init = nullptr; v.init_t = false; }
        else { v.init = nullptr; v.init_t = false; }
      }
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "Transpose") {
      if (!need(0)) return fail("missing input");
      const Shape& s = in[0]->shape;
      std::vector<int64_t> perm;
      if (const OnnxAttr* a = n.attr("perm")) perm = a->ints;
      if (perm.empty()) for (int i = static_cast<int>(s.size()) - 1; i >= 0; --i) perm.push_back(i);
      if (perm.size() != s.size()) return fail("Transpose: perm rank mismatch");
      {
        std::set<int64_t> seen;
        for (int64_t p : perm)
          if (p < 0 || p >= static_cast<int64_t>(s.size()) || !seen.insert(p).second) return fail("Transpose: perm is not a permutation");
      }
      AVal v;
      v.dtype = in[0]->dtype;
      v.tainted = in[0]->tainted;
      for (int64_t p : perm) v.shape.push_back(s[p]);
      if (!v.tainted) {
        if (!in[0]->has_data && in[0]->init != nullptr && s.size() == 2 && perm[0] == 1 && perm[1] == 0 &&
            in[0]->init->dims.size() == 2 && in[0]->init->dims == (in[0]->init_t ? v.shape : s)) {
          v.init = in[0]->init;
          v.init_name = in[0]->init_name;
          v.init_t = !in[0]->init_t;
        } else if (numel(s) <= kMaxConst && materialize(in[0])) {
          const Shape st = strides_of(s);
          Shape pst(s.size());
          for (size_t i = 0; i < perm.size(); ++i) pst[i] = st[perm[i]];
          v.data.resize(in[0]->data.size());
          for_each(v.shape, [&](int64_t i, const Shape& ix) { v.data[i] = in[0]->data[dot(ix, pst)]; });
          v.has_data = true;
        }
      }
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "Concat") {
      int64_t axis = n.attr_i("axis", 0);
      std::vector<AVal*> xs;
      for (AVal* v : in) if (v != nullptr) xs.push_back(v);
      if (xs.empty()) return fail("Concat without inputs");
      const size_t r = xs[0]->shape.size();
      if (axis < 0) axis += static_cast<int64_t>(r);
      if (axis < 0 || axis >= static_cast<int64_t>(r)) return fail("Concat: axis out of range");
      AVal v;
      v.dtype = xs[0]->dtype;
      v.shape = xs[0]->shape;
      v.shape[axis] = 0;
      bool all_data = true;
      for (AVal* x : xs) {
        if (x->shape.size() != r) return fail("Concat: rank mismatch");
        for (size_t d = 0; d < r; ++d)
          if (static_cast<int64_t>(d) != axis && x->shape[d] != xs[0]->shape[d]) return fail("Concat: shapes differ off the axis");
        v.shape[axis] += x->shape[axis];
        v.tainted |= x->tainted;
      }
      if (!v.tainted && numel(v.shape) <= kMaxConst) {
        for (AVal* x : xs) all_data &= materialize(x);
        if (all_data) {
          v.data.resize(static_cast<size_t>(numel(v.shape)));
          const Shape ost = strides_of(v.shape);
          int64_t off = 0;
          for (AVal* x : xs) {
            for_each(x->shape, [&](int64_t i, const Shape& ix) {
              int64_t o = dot(ix, ost) + off * ost[axis];
              v.data[o] = x->data[i];
            });
            off += x->shape[axis];
          }
          v.has_data = true;
        }
      }
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "Slice") {
      if (!need(0)) return fail("missing input");
      const Shape& s = in[0]->shape;
      std::vector<int64_t> starts, ends, axes, steps;
      if (need(1)) {
        if (!ints_of(in[1], &starts) || !need(2) || !ints_of(in[2], &ends)) return fail("Slice bounds are not constants");
        if (need(3) && !ints_of(in[3], &axes)) return fail("Slice axes are not constants");
        if (need(4) && !ints_of(in[4], &steps)) return fail("Slice steps are not constants");
      } else {
        if (const OnnxAttr* a = n.attr("starts")) starts = a->ints;
        if (const OnnxAttr* a = n.attr("ends")) ends = a->ints;
        if (const OnnxAttr* a = n.attr("axes")) axes = a->ints;
      }
      if (axes.empty()) for (size_t i = 0; i < starts.size(); ++i) axes.push_back(static_cast<int64_t>(i));
      if (steps.empty()) steps.assign(starts.size(), 1);
      Shape b(s.size(), 0), st(s.size(), 1);
      AVal v;
      v.dtype = in[0]->dtype;
      v.tainted = in[0]->tainted;
      v.shape = s;
      for (size_t k = 0; k < starts.size(); ++k) {
        int64_t a = axes[k] < 0 ? axes[k] + static_cast<int64_t>(s.size()) : axes[k];
        if (a < 0 || a >= static_cast<int64_t>(s.size())) return fail("Slice: axis out of range");
        if (steps[k] <= 0) return fail("Slice: non-positive step");
        const int64_t d = s[a];
        int64_t lo = starts[k] < 0 ? starts[k] + d : starts[k];
        int64_t hi = ends[k] < 0 ? ends[k] + d : ends[k];
        lo = std::max<int64_t>(0, std::min(lo, d));
        hi = std::max<int64_t>(0, std::min(hi, d));
        b[a] = lo;
        st[a] = steps[k];
        v.shape[a] = hi > lo ? (hi - lo + steps[k] - 1) / steps[k] : 0;
      }
      if (!v.tainted && numel(v.shape) <= kMaxConst && materialize(in[0])) {
        const Shape ist = strides_of(s);
        v.data.resize(static_cast<size_t>(numel(v.shape)));
        for_each(v.shape, [&](int64_t i, const Shape& ix) {
          int64_t o = 0;
          for (size_t d = 0; d < ix.size(); ++d) o += (b[d] + ix[d] * st[d]) * ist[d];
          v.data[i] = in[0]->data[o];
        });
        v.has_data = true;
      }
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "Gather") {
      if (!need(0) || !need(1)) return fail("missing input");
      const Shape& s = in[0]->shape;
      int64_t axis = n.attr_i("axis", 0);
      if (axis < 0) axis += static_cast<int64_t>(s.size());
      if (axis < 0 || axis >= static_cast<int64_t>(s.size())) return fail("Gather: axis out of range");
      AVal v;
      v.dtype = in[0]->dtype;
      v.tainted = in[0]->tainted || in[1]->tainted;
      for (int64_t i = 0; i < axis; ++i) v.shape.push_back(s[i]);
      for (int64_t d : in[1]->shape) v.shape.push_back(d);
      for (size_t i = axis + 1; i < s.size(); ++i) v.shape.push_back(s[i]);
      if (!v.tainted && numel(v.shape) <= kMaxConst && numel(s) <= kMaxConst && materialize(in[1]) && materialize(in[0])) {
        const Shape ist = strides_of(s);
        const size_t ir = in[1]->shape.size();
        const Shape jst = strides_of(in[1]->shape);
        v.data.resize(static_cast<size_t>(numel(v.shape)));
        bool oob = false;
        for_each(v.shape, [&](int64_t i, const Shape& ix) {
          int64_t j = 0;
          for (size_t d = 0; d < ir; ++d) j += ix[axis + d] * jst[d];
          int64_t g = to_i64(in[1]->data[j]);
          if (g < 0) g += s[axis];
          if (g < 0 || g >= s[axis]) { oob = true; return; }
          int64_t o = g * ist[axis];
          for (int64_t d = 0; d < axis; ++d) o += ix[d] * ist[d];
          for (size_t d = axis + 1; d < s.size(); ++d) o += ix[d + ir - 1] * ist[d];
          v.data[i] = in[0]->data[o];
        });
        if (oob) return fail("Gather: index out of range");
        v.has_data = true;
      }
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "Expand") {
      if (!need(0) || !need(1)) return fail("missing input");
      std::vector<int64_t> t;
      if (!ints_of(in[1], &t)) return fail("Expand shape is not a constant");
      AVal v;
      v.dtype = in[0]->dtype;
      v.tainted = in[0]->tainted;
      if (!bshape(in[0]->shape, t, &v.shape)) return fail("Expand: shapes do not broadcast");
      if (!v.tainted) {
        if (numel(v.shape) == numel(in[0]->shape) && !in[0]->has_data && in[0]->init != nullptr && !in[0]->init_t) {
          v.init = in[0]->init;
          v.init_name = in[0]->init_name;
        } else if (numel(v.shape) <= kMaxConst && materialize(in[0])) {
          const Shape st = bstrides(in[0]->shape, v.shape);
          v.data.resize(static_cast<size_t>(numel(v.shape)));
          for_each(v.shape, [&](int64_t i, const Shape& ix) { v.data[i] = in[0]->data[dot(ix, st)]; });
          v.has_data = true;
        }
      }
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "ConstantOfShape") {
      std::vector<int64_t> t;
      if (!need(0) || !ints_of(in[0], &t)) return fail("ConstantOfShape shape is not a constant");
      AVal v;
      double val = 0.0;
      v.dtype = 1;
      if (const OnnxAttr* a = n.attr("value")) {
        if (!a->t) return fail("ConstantOfShape value must be inline");
        std::vector<double> d;
        if (!read_tensor(*a->t, &d) || d.empty()) return fail("ConstantOfShape value unreadable");
        val = d[0];
        v.dtype = a->t->data_type;
      }
      v.shape = t;
      if (!shape_ok(t)) return fail("ConstantOfShape: invalid shape");
      if (numel(t) <= kMaxConst) {
        v.data.assign(static_cast<size_t>(numel(t)), val);
        v.has_data = true;
      }
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "Range") {
      if (!need(0) || !need(1) || !need(2)) return fail("missing input");
      if (any_taint) return fail("Range over input-dependent bounds");
      if (!materialize(in[0]) || !materialize(in[1]) || !materialize(in[2])) return fail("Range bounds are not constants");
      const double a = in[0]->data[0], b = in[1]->data[0], d = in[2]->data[0];
      if (d == 0) return fail("Range: zero step");
      AVal v;
      v.dtype = in[0]->dtype;
      const double span = ceil((b - a) / d);
      if (!(span < static_cast<double>(kMaxConst))) return fail("Range: too many elements");
      const int64_t cnt = std::max<int64_t>(0, static_cast<int64_t>(span));
      for (int64_t i = 0; i < cnt; ++i) v.data.push_back(a + i * d);
      v.shape = {cnt};
      v.has_data = true;
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "Trilu") {
      if (!need(0)) return fail("missing input");
      AVal v = *in[0];
      v.init = nullptr;
      v.init_t = false;
      if (!v.tainted) {
        AVal src = *in[0];
        int64_t k = 0;
        if (need(1)) {
          std::vector<int64_t> kk;
          if (!ints_of(in[1], &kk) || kk.empty()) return fail("Trilu k is not a constant");
          k = kk[0];
        }
        const bool upper = n.attr_i("upper", 1) != 0;
        if (src.shape.size() >= 2 && numel(src.shape) <= kMaxConst && materialize(&src)) {
          v.data = src.data;
          const size_t r = src.shape.size();
          for_each(src.shape, [&](int64_t i, const Shape& ix) {
            const int64_t row = ix[r - 2], col = ix[r - 1];
            const bool keep = upper ? (col - row >= k) : (col - row <= k);
            if (!keep) v.data[i] = 0.0;
          });
          v.has_data = true;
        } else {
          v.has_data = false;
          v.data.clear();
        }
      }
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "Where") {
      if (!need(0) || !need(1) || !need(2)) return fail("missing input");
      AVal v;
      Shape t;
      if (!bshape(in[0]->shape, in[1]->shape, &t) || !bshape(t, in[2]->shape, &v.shape)) return fail("Where: shapes do not broadcast");
      v.dtype = in[1]->dtype;
      v.tainted = any_taint;
      if (!v.tainted && numel(v.shape) <= kMaxConst && materialize(in[0]) && materialize(in[1]) && materialize(in[2])) {
        const Shape s0 = bstrides(in[0]->shape, v.shape), s1 = bstrides(in[1]->shape, v.shape), s2 = bstrides(in[2]->shape, v.shape);
        v.data.resize(static_cast<size_t>(numel(v.shape)));
        for_each(v.shape, [&](int64_t i, const Shape& ix) {
          v.data[i] = in[0]->data[dot(ix, s0)] != 0.0 ? in[1]->data[dot(ix, s1)] : in[2]->data[dot(ix, s2)];
        });
        v.has_data = true;
      }
      put(n, idx, std::move(v));
      return true;
    }
    static const std::set<std::string> kBinary = {"Add", "Sub", "Mul", "Div", "Pow", "Mod", "Equal", "Less", "Greater",
                                                  "LessOrEqual", "GreaterOrEqual", "And", "Or", "Max", "Min"};
    if (kBinary.count(op)) {
      if (!need(0) || !need(1)) return fail("missing input");
      AVal v;
      if (!bshape(in[0]->shape, in[1]->shape, &v.shape))
        return fail(op + ": shapes " + shape_str(in[0]->shape) + " and " + shape_str(in[1]->shape) + " do not broadcast");
      const bool cmp = op == "Equal" || op == "Less" || op == "Greater" || op == "LessOrEqual" || op == "GreaterOrEqual";
      v.dtype = cmp ? 9 : (is_float_dt(in[0]->dtype) ? in[0]->dtype : in[1]->dtype);
      v.tainted = any_taint;
      if (!v.tainted && numel(v.shape) <= kMaxConst && materialize(in[0]) && materialize(in[1])) {
        const Shape s0 = bstrides(in[0]->shape, v.shape), s1 = bstrides(in[1]->shape, v.shape);
        const bool integer = !is_float_dt(in[0]->dtype) && !is_float_dt(in[1]->dtype);
        const bool fmod_attr = n.attr_i("fmod", 0) != 0;
        v.data.resize(static_cast<size_t>(numel(v.shape)));
        for_each(v.shape, [&](int64_t i, const Shape& ix) {
          const double a = in[0]->data[dot(ix, s0)], b = in[1]->data[dot(ix, s1)];
          double r = 0.0;
          if (op == "Add") r = a + b;
          else if (op == "Sub") r = a - b;
          else if (op == "Mul") r = a * b;
          else if (op == "Div") r = integer ? (b == 0 ? 0.0 : trunc(a / b)) : a / b;
          else if (op == "Pow") r = pow(a, b);
          else if (op == "Mod") {
            if (b == 0) r = 0.0;
            else if (fmod_attr) r = fmod(a, b);
            else { r = fmod(a, b); if (r != 0.0 && ((r < 0) != (b < 0))) r += b; }
          } else if (op == "Equal") r = a == b;
          else if (op == "Less") r = a < b;
          else if (op == "Greater") r = a > b;
          else if (op == "LessOrEqual") r = a <= b;
          else if (op == "GreaterOrEqual") r = a >= b;
          else if (op == "And") r = (a != 0.0) && (b != 0.0);
          else if (op == "Or") r = (a != 0.0) || (b != 0.0);
          else if (op == "Max") r = std::max(a, b);
          else if (op == "Min") r = std::min(a, b);
          v.data[i] = r;
        });
        v.has_data = true;
      }
      put(n, idx, std::move(v));
      return true;
    }
    static const std::set<std::string> kUnary = {"Sqrt", "Neg", "Not", "Abs", "Floor", "Ceil", "Exp", "Log", "Erf", "Tanh",
                                                 "Sigmoid", "Relu", "Reciprocal", "Softmax", "LayerNormalization", "Clip",
                                                 "Gelu", "LogSoftmax", "Softplus", "HardSigmoid", "HardSwish"};
    if (kUnary.count(op)) {
      if (!need(0)) return fail("missing input");
      AVal v;
      v.shape = in[0]->shape;
      v.dtype = in[0]->dtype;
      v.tainted = any_taint;
      const bool pointwise = op != "Softmax" && op != "LayerNormalization" && op != "Clip" && op != "LogSoftmax" &&
                             op != "Gelu" && op != "Softplus" && op != "HardSigmoid" && op != "HardSwish";
      if (!v.tainted && pointwise && numel(v.shape) <= kMaxConst && materialize(in[0])) {
        v.data = in[0]->data;
        for (double& d : v.data) {
          if (op == "Sqrt") d = sqrt(d);
          else if (op == "Neg") d = -d;
          else if (op == "Not") d = d == 0.0;
          else if (op == "Abs") d = fabs(d);
          else if (op == "Floor") d = floor(d);
          else if (op == "Ceil") d = ceil(d);
          else if (op == "Exp") d = exp(d);
          else if (op == "Log") d = log(d);
          else if (op == "Erf") d = erf(d);
          else if (op == "Tanh") d = tanh(d);
          else if (op == "Sigmoid") d = 1.0 / (1.0 + exp(-d));
          else if (op == "Relu") d = d > 0 ? d : 0.0;
          else if (op == "Reciprocal") d = 1.0 / d;
        }
        v.has_data = true;
      }
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "MatMul") {
      if (!need(0) || !need(1)) return fail("missing input");
      const Shape &a = in[0]->shape, &b = in[1]->shape;
      if (a.size() < 2 || b.size() < 2) return fail("MatMul with a 1-D operand");
      const int64_t M = a[a.size() - 2], K = a[a.size() - 1], N = b[b.size() - 1];
      if (b[b.size() - 2] != K) return fail("MatMul " + shape_str(a) + " x " + shape_str(b) + ": inner dimensions differ");
      Shape ba(a.begin(), a.end() - 2), bb(b.begin(), b.end() - 2), bo;
      if (!bshape(ba, bb, &bo)) return fail("MatMul: batch dimensions do not broadcast");
      AVal v;
      v.dtype = in[0]->dtype;
      v.tainted = any_taint;
      v.shape = bo;
      v.shape.push_back(M);
      v.shape.push_back(N);
      const int64_t batch = numel(bo);
      if (!v.tainted && numel(v.shape) <= kMaxConst && batch * M * N * K <= (int64_t(1) << 28) && materialize(in[0]) &&
          materialize(in[1])) {
        const Shape sa = bstrides(ba, bo), sb = bstrides(bb, bo);
        v.data.assign(static_cast<size_t>(numel(v.shape)), 0.0);
        for_each(bo, [&](int64_t bi, const Shape& ix) {
          const double* pa = in[0]->data.data() + dot(ix, sa) * M * K;
          const double* pb = in[1]->data.data() + dot(ix, sb) * K * N;
          double* po = v.data.data() + bi * M * N;
          for (int64_t i = 0; i < M; ++i)
            for (int64_t k = 0; k < K; ++k) {
              const double x = pa[i * K + k];
              for (int64_t j = 0; j < N; ++j) po[i * N + j] += x * pb[k * N + j];
            }
        });
        v.has_data = true;
      }
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "Gemm") {
      if (!need(0) || !need(1)) return fail("missing input");
      const Shape &a = in[0]->shape, &b = in[1]->shape;
      if (a.size() != 2 || b.size() != 2) return fail("Gemm operands must be 2-D");
      const bool ta = n.attr_i("transA", 0) != 0, tb = n.attr_i("transB", 0) != 0;
      const int64_t M = ta ? a[1] : a[0], K = ta ? a[0] : a[1], Kb = tb ? b[1] : b[0], N = tb ? b[0] : b[1];
      if (K != Kb) return fail("Gemm: inner dimensions differ");
      AVal v;
      v.dtype = in[0]->dtype;
      v.tainted = any_taint;
      v.shape = {M, N};
      if (need(2)) {
        Shape chk;
        if (in[2]->shape.size() > 2 || !bshape(in[2]->shape, v.shape, &chk) || chk != v.shape) return fail("Gemm: C does not broadcast to [M, N]");
      }
      if (!v.tainted && M * N <= kMaxConst && M * N * K <= (int64_t(1) << 28) && materialize(in[0]) && materialize(in[1]) &&
          (!need(2) || materialize(in[2]))) {
        const double alpha = n.attr_f("alpha", 1.f), beta = n.attr_f("beta", 1.f);
        v.data.assign(static_cast<size_t>(M * N), 0.0);
        for (int64_t i = 0; i < M; ++i)
          for (int64_t j = 0; j < N; ++j) {
            double acc = 0.0;
            for (int64_t k = 0; k < K; ++k)
              acc += in[0]->data[ta ? k * M + i : i * K + k] * in[1]->data[tb ? j * K + k : k * N + j];
            v.data[i * N + j] = alpha * acc;
          }
        if (need(2)) {
          const Shape sc = bstrides(in[2]->shape, v.shape);
          for_each(v.shape, [&](int64_t i, const Shape& ix) { v.data[i] += beta * in[2]->data[dot(ix, sc)]; });
        }
        v.has_data = true;
      }
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "Conv") {
      if (!need(0) || !need(1)) return fail("missing input");
      const Shape &x = in[0]->shape, &w = in[1]->shape;
      if (x.size() != 4 || w.size() != 4) return fail("Conv: only 2-D convolutions are analysed");
      std::vector<int64_t> strides = {1, 1}, pads = {0, 0, 0, 0}, dil = {1, 1};
      if (const OnnxAttr* a = n.attr("strides")) strides = a->ints;
      if (const OnnxAttr* a = n.attr("pads")) pads = a->ints;
      if (const OnnxAttr* a = n.attr("dilations")) dil = a->ints;
      if (strides.size() != 2 || pads.size() != 4 || dil.size() != 2) return fail("Conv: malformed attributes");
      for (int64_t v : strides) if (v <= 0 || v > 4096) return fail("Conv: invalid stride");
      for (int64_t v : dil) if (v <= 0 || v > 4096) return fail("Conv: invalid dilation");
      for (int64_t v : pads) if (v < 0 || v > 4096) return fail("Conv: invalid padding");
      AVal v;
      v.dtype = in[0]->dtype;
      v.tainted = any_taint;
      v.shape = {x[0], w[0], (x[2] + pads[0] + pads[2] - dil[0] * (w[2] - 1) - 1) / strides[0] + 1,
                 (x[3] + pads[1] + pads[3] - dil[1] * (w[3] - 1) - 1) / strides[1] + 1};
      if (v.shape[2] <= 0 || v.shape[3] <= 0) return fail("Conv: kernel larger than the padded input");
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "ReduceL2" || op == "ReduceMean" || op == "ReduceSum" || op == "ReduceMax" || op == "ReduceMin" ||
        op == "ReduceProd" || op == "ArgMax" || op == "ArgMin") {
      if (!need(0)) return fail("missing input");
      const Shape& s = in[0]->shape;
      std::vector<int64_t> axes;
      bool present = false;
      if (op == "ArgMax" || op == "ArgMin") {
        axes = {n.attr_i("axis", 0)};
        present = true;
      } else if (!axes_arg(n, 1, &axes, &present)) return false;
      const bool keep = n.attr_i("keepdims", 1) != 0;
      std::set<int64_t> ax;
      if (!present || axes.empty()) {
        if (n.attr_i("noop_with_empty_axes", 0) != 0) { AVal v = *in[0]; v.init = nullptr; put(n, idx, std::move(v)); return true; }
        for (size_t i = 0; i < s.size(); ++i) ax.insert(static_cast<int64_t>(i));
      }
      for (int64_t a : axes) {
        const int64_t p = a < 0 ? a + static_cast<int64_t>(s.size()) : a;
        if (p < 0 || p >= static_cast<int64_t>(s.size())) return fail(op + ": axis out of range");
        ax.insert(p);
      }
      AVal v;
      v.dtype = (op == "ArgMax" || op == "ArgMin") ? 7 : in[0]->dtype;
      v.tainted = any_taint;
      for (size_t i = 0; i < s.size(); ++i) {
        if (!ax.count(static_cast<int64_t>(i))) v.shape.push_back(s[i]);
        else if (keep) v.shape.push_back(1);
      }
      if (!v.tainted && op == "ReduceProd" && materialize(in[0]) && ax.size() == s.size()) {
        double p = 1.0;
        for (double d : in[0]->data) p *= d;
        v.data = {p};
        v.has_data = true;
      }
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "GlobalAveragePool") {
      if (!need(0)) return fail("missing input");
      AVal v;
      v.dtype = in[0]->dtype;
      v.tainted = any_taint;
      v.shape = in[0]->shape;
      for (size_t i = 2; i < v.shape.size(); ++i) v.shape[i] = 1;
      put(n, idx, std::move(v));
      return true;
    }
    if (op == "Split") {
      if (!need(0)) return fail("missing input");
      const Shape& s = in[0]->shape;
      int64_t axis = n.attr_i("axis", 0);
      if (axis < 0) axis += static_cast<int64_t>(s.size());
      std::vector<int64_t> parts;
      if (need(1)) {
        if (!ints_of(in[1], &parts)) return fail("Split sizes are not constants");
      } else if (const OnnxAttr* a = n.attr("split")) {
        parts = a->ints;
      } else {
        const int64_t k = n.attr_i("num_outputs", static_cast<int64_t>(n.outputs.size()));
        if (k <= 0 || k > 1024 || axis < 0 || axis >= static_cast<int64_t>(s.size())) return fail("Split: invalid num_outputs / axis");
        const int64_t chunk = (s[axis] + k - 1) / k;
        for (int64_t i = 0; i < k; ++i) parts.push_back(std::min(chunk, s[axis] - i * chunk));
      }
      if (axis < 0 || axis >= static_cast<int64_t>(s.size())) return fail("Split: axis out of range");
      {
        int64_t total = 0;
        for (int64_t p : parts) {
          if (p < 0) return fail("Split: negative part");
          total += p;
        }
        if (total != s[axis] || parts.size() > 1024) return fail("Split: parts do not cover the axis");
      }
      int64_t off = 0;
      const bool can = !in[0]->tainted && numel(s) <= kMaxConst && materialize(in[0]);
      const Shape ist = strides_of(s);
      for (size_t k = 0; k < parts.size(); ++k) {
        AVal v;
        v.dtype = in[0]->dtype;
        v.tainted = in[0]->tainted;
        v.shape = s;
        v.shape[axis] = parts[k];
        if (can) {
          v.data.resize(static_cast<size_t>(numel(v.shape)));
          for_each(v.shape, [&](int64_t i, const Shape& ix) {
            int64_t o = off * ist[axis];
            for (size_t d = 0; d < ix.size(); ++d) o += ix[d] * ist[d];
            v.data[i] = in[0]->data[o];
          });
          v.has_data = true;
        }
        put(n, idx, std::move(v), k);
        off += parts[k];
      }
      return true;
    }
    return fail("operator '" + op + "' is not supported by the graph analyser");
  }
};

// ----------------------------------------------------------------------------------------------------------
// 2. sites and tokens
// ----------------------------------------------------------------------------------------------------------
struct LinearSite {
  int node = -1;
  int kind = 0;  // 0 MatMul, 1 Gemm, 2 Conv
  std::string data_in, raw_out, out;
  AVal* w = nullptr;
  bool w_kn = false;  // weight bytes are [K, N] row-major (else [N, K])
  AVal* bias = nullptr;
  int bias_node = -1;
  int64_t N = 0, K = 0;
};

struct AttnSite {
  int softmax_node = -1;
  int64_t heads = 0, Tq = 0, Tk = 0, hd = 0;
  double scale = 1.0;  // product of the scalar factors between the projections and the softmax input
  int mask = 0;  // 0 none, 1 causal
  int q_site = -1, k_site = -1, v_site = -1;
  int q_sel = -1, k_sel = -1, v_sel = -1;  // chunk index inside a fused projection (-1 unknown / not fused)
  bool q_const = false;
  AVal* q_val = nullptr;  // value entering QK^T on the query side when it does not depend on the input
};

enum TokKind { TK_CONV, TK_EMBED, TK_CLS, TK_POS, TK_LN, TK_LIN, TK_SOFTMAX, TK_ACT, TK_SEL, TK_L2 };
struct Token {
  TokKind kind;
  int node = -1;
  int ref = -1;      // site index (TK_LIN / TK_CONV), attention index (TK_SOFTMAX), activation id (TK_ACT)
  AVal* cval = nullptr;  // constant operand (class token, positional table, embedding table)
  int sel = 0;       // TK_SEL: 0 first, 1 last, 2 argmax
};

const char* tok_name(TokKind k) {
  static const char* names[] = {"conv", "embed", "cls", "pos", "ln", "linear", "softmax", "act", "select", "l2norm"};
  return names[k];
}

struct Recognizer {
  OnnxModel& m;
  Analyzer an;
  std::unordered_map<std::string, int> producer;
  std::unordered_map<std::string, std::vector<int>> consumers;
  std::vector<LinearSite> sites;
  std::unordered_map<std::string, int> site_of_output;
  std::vector<AttnSite> attns;
  std::set<int> mask_add_nodes, bias_add_nodes;
  std::vector<Token> toks;
  std::string err;
  // results
  std::map<std::string, OnnxTensor> out_tensors;
  std::map<std::string, std::string> out_meta;
  std::vector<GraphBinding> bindings;

  explicit Recognizer(OnnxModel& model) : m(model), an(model) {}

  bool fail(const std::string& s) {
    if (err.empty()) err = s;
    return false;
  }

  static bool is_view_op(const std::string& op) {
    return op == "Reshape" || op == "Transpose" || op == "Squeeze" || op == "Unsqueeze" || op == "Identity" ||
           op == "Cast" || op == "Flatten" || op == "Expand" || op == "Dropout";
  }
  bool scalar_value(AVal* v, double* out) {
    if (v == nullptr || v->tainted || numel(v->shape) != 1 || !materialize(v)) return false;
    *out = v->data[0];
    return true;
  }

  void index_graph() {
    for (size_t i = 0; i < m.nodes.size(); ++i) {
      for (const std::string& o : m.nodes[i].outputs) if (!o.empty()) producer[o] = static_cast<int>(i);
      for (const std::string& s : m.nodes[i].inputs) if (!s.empty()) consumers[s].push_back(static_cast<int>(i));
    }
  }

  static bool weight_like(const AVal* v, size_t rank) {
    return v != nullptr && !v->tainted && v->init != nullptr && v->shape.size() == rank &&
           is_float_dt(v->dtype) && v->init->dims.size() == rank;
  }

  void find_bias(LinearSite* s) {
    s->out = s->raw_out;
    auto it = consumers.find(s->raw_out);
    if (it == consumers.end()) return;
    std::vector<int> users;
    for (int u : it->second) if (m.nodes[u].op_type != "Shape") users.push_back(u);
    if (users.size() != 1) return;
    const OnnxNode& c = m.nodes[users[0]];
    if (c.op_type != "Add" || c.inputs.size() != 2) return;
    const std::string& other = c.inputs[0] == s->raw_out ? c.inputs[1] : c.inputs[0];
    AVal* b = an.get(other);
    if (b == nullptr || b->tainted || numel(b->shape) != s->N || b->shape.empty() || b->shape.back() != s->N) return;
    s->bias = b;
    s->bias_node = users[0];
    s->out = c.outputs[0];
    bias_add_nodes.insert(users[0]);
  }

  bool find_sites() {
    for (size_t i = 0; i < m.nodes.size(); ++i) {
      const OnnxNode& n = m.nodes[i];
      if (n.outputs.empty()) continue;
      AVal* o = an.get(n.outputs[0]);
      if (o == nullptr || !o->tainted) continue;
      LinearSite s;
      s.node = static_cast<int>(i);
      if (n.op_type == "MatMul") {
        AVal *a = an.get(n.inputs[0]), *b = an.get(n.inputs[1]);
        if (!(a && a->tainted && weight_like(b, 2))) continue;
        s.kind = 0;
        s.data_in = n.inputs[0];
        s.w = b;
        s.K = b->shape[0];
        s.N = b->shape[1];
        s.w_kn = !b->init_t;
        s.raw_out = n.outputs[0];
        find_bias(&s);
      } else if (n.op_type == "Gemm") {
        AVal *a = an.get(n.inputs[0]), *b = an.get(n.inputs[1]);
        if (!(a && a->tainted && weight_like(b, 2))) continue;
        if (n.attr_i("transA", 0) != 0 || n.attr_f("alpha", 1.f) != 1.f || n.attr_f("beta", 1.f) != 1.f)
          return fail("Gemm with transA / alpha / beta is not supported");
        const bool tb = n.attr_i("transB", 0) != 0;
        s.kind = 1;
        s.data_in = n.inputs[0];
        s.w = b;
        s.K = tb ? b->shape[1] : b->shape[0];
        s.N = tb ? b->shape[0] : b->shape[1];
        s.w_kn = tb == b->init_t;  // !tb & !t -> [K,N]; tb & !t -> [N,K]; a Transpose in front flips it
        s.raw_out = s.out = n.outputs[0];
        if (n.inputs.size() > 2 && !n.inputs[2].empty()) {
          AVal* c = an.get(n.inputs[2]);
          if (c == nullptr || c->tainted || numel(c->shape) != s.N) return fail("Gemm bias is not a constant vector");
          s.bias = c;
        } else {
          find_bias(&s);
        }
      } else if (n.op_type == "Conv") {
        AVal *a = an.get(n.inputs[0]), *b = an.get(n.inputs[1]);
        if (!(a && a->tainted && weight_like(b, 4)) || b->init_t) continue;
        s.kind = 2;
        s.data_in = n.inputs[0];
        s.w = b;
        s.N = b->shape[0];
        s.K = b->shape[1] * b->shape[2] * b->shape[3];
        s.w_kn = false;
        s.raw_out = s.out = n.outputs[0];
        if (n.inputs.size() > 2 && !n.inputs[2].empty()) {
          AVal* c = an.get(n.inputs[2]);
          if (c == nullptr || c->tainted || numel(c->shape) != s.N) return fail("Conv bias is not a constant vector");
          s.bias = c;
        }
      } else {
        continue;
      }
      const int si = static_cast<int>(sites.size());
      site_of_output[s.raw_out] = si;
      site_of_output[s.out] = si;
      sites.push_back(s);
    }
    return true;
  }

  // Walks from `name` towards the producers through view ops, constant-index selects and scalar Mul/Div until a
  // linear-site output (or an input-independent value when `allow_const`) is reached.
  bool trace_operand(std::string name, int* site, int* sel, double* scale, AVal** const_val) {
    *site = -1;
    *sel = -1;
    for (int guard = 0; guard < 64; ++guard) {
      auto so = site_of_output.find(name);
      if (so != site_of_output.end()) {
        *site = so->second;
        return true;
      }
      AVal* v = an.get(name);
      if (v == nullptr) return fail("attention operand '" + name + "' is undefined");
      if (!v->tainted) {
        if (const_val != nullptr) *const_val = v;
        return true;
      }
      auto p = producer.find(name);
      if (p == producer.end()) return fail("attention operand is fed directly by a graph input");
      const OnnxNode& n = m.nodes[p->second];
      if (is_view_op(n.op_type)) {
        name = n.inputs[0];
      } else if (n.op_type == "Mul" || n.op_type == "Div") {
        double c;
        AVal *a = an.get(n.inputs[0]), *b = an.get(n.inputs[1]);
        if (scalar_value(b, &c)) {
          *scale *= n.op_type == "Mul" ? c : 1.0 / c;
          name = n.inputs[0];
        } else if (n.op_type == "Mul" && scalar_value(a, &c)) {
          *scale *= c;
          name = n.inputs[1];
        } else return fail("attention operand is scaled by a non-scalar");
      } else if (n.op_type == "Gather") {
        AVal* ix = an.get(n.inputs[1]);
        std::vector<int64_t> iv;
        if (ix == nullptr || ix->tainted || !an.ints_of(ix, &iv) || iv.size() != 1)
          return fail("attention operand is gathered with a non-constant index");
        *sel = static_cast<int>(iv[0]);
        name = n.inputs[0];
      } else if (n.op_type == "Split") {
        for (size_t k = 0; k < n.outputs.size(); ++k) if (n.outputs[k] == name) *sel = static_cast<int>(k);
        name = n.inputs[0];
      } else if (n.op_type == "Slice") {
        // chunk index = start / length along the sliced axis
        AVal* in0 = an.get(n.inputs[0]);
        std::vector<int64_t> st, ax;
        if (n.inputs.size() < 3 || !an.ints_of(an.get(n.inputs[1]), &st) || st.size() != 1)
          return fail("attention operand is sliced with non-constant bounds");
        int64_t axis = 0;
        if (n.inputs.size() > 3 && an.ints_of(an.get(n.inputs[3]), &ax) && ax.size() == 1) axis = ax[0];
        if (axis < 0) axis += static_cast<int64_t>(in0->shape.size());
        if (axis < 0 || axis >= static_cast<int64_t>(in0->shape.size()) || v->shape.size() != in0->shape.size())
          return fail("attention operand slice axis out of range");
        const int64_t len = v->shape[axis];
        int64_t s0 = st[0] < 0 ? st[0] + in0->shape[axis] : st[0];
        if (len <= 0 || s0 % len != 0) return fail("attention operand slice is not chunk aligned");
        *sel = static_cast<int>(s0 / len);
        name = n.inputs[0];
      } else {
        return fail("unexpected operator '" + n.op_type + "' between a projection and the attention product");
      }
    }
    return fail("attention operand trace did not terminate");
  }

  bool analyse_softmax(int node_idx) {
    const OnnxNode& sm = m.nodes[node_idx];
    AVal* x = an.get(sm.inputs[0]);
    if (x == nullptr || !x->tainted) return true;  // constant softmax: not an attention site
    if (x->shape.size() < 2) return fail("Softmax over a rank-1 tensor");
    const int64_t axis = sm.attr_i("axis", -1);
    if (axis != -1 && axis != static_cast<int64_t>(x->shape.size()) - 1) return fail("Softmax is not over the last axis");
    AttnSite a;
    a.softmax_node = node_idx;
    a.Tq = x->shape[x->shape.size() - 2];
    a.Tk = x->shape[x->shape.size() - 1];
    if (a.Tq <= 0 || a.Tk <= 0) return fail("Softmax over an empty tensor");
    a.heads = numel(x->shape) / (a.Tq * a.Tk) / (an.batch0 > 0 ? an.batch0 : 1);
    // back from the softmax input to QK^T
    std::string cur = sm.inputs[0];
    int qk_node = -1;
    AVal* mask = nullptr;
    for (int guard = 0; guard < 32 && qk_node < 0; ++guard) {
      auto p = producer.find(cur);
      if (p == producer.end()) return fail("Softmax input has no producer");
      const OnnxNode& n = m.nodes[p->second];
      if (n.op_type == "MatMul") {
        qk_node = p->second;
      } else if (n.op_type == "Add") {
        AVal *u = an.get(n.inputs[0]), *w = an.get(n.inputs[1]);
        if (u->tainted && !w->tainted) { mask = w; cur = n.inputs[0]; }
        else if (!u->tainted && w->tainted) { mask = u; cur = n.inputs[1]; }
        else return fail("attention scores are added to an input-dependent tensor (attention_mask inputs are not supported)");
        mask_add_nodes.insert(p->second);
      } else if (n.op_type == "Mul" || n.op_type == "Div") {
        double c;
        if (scalar_value(an.get(n.inputs[1]), &c)) { a.scale *= n.op_type == "Mul" ? c : 1.0 / c; cur = n.inputs[0]; }
        else if (n.op_type == "Mul" && scalar_value(an.get(n.inputs[0]), &c)) { a.scale *= c; cur = n.inputs[1]; }
        else return fail("attention scores are scaled by a non-scalar");
      } else if (is_view_op(n.op_type)) {
        cur = n.inputs[0];
      } else {
        return fail("unexpected operator '" + n.op_type + "' between QK^T and Softmax");
      }
    }
    if (qk_node < 0) return fail("no QK^T MatMul found in front of Softmax");
    const OnnxNode& qk = m.nodes[qk_node];
    AVal* qa = an.get(qk.inputs[0]);
    a.hd = qa->shape.back();
    if (a.hd <= 0) return fail("attention with an empty head dimension");
    double qs = 1.0, ks = 1.0, vs = 1.0;
    AVal* qconst = nullptr;
    if (!trace_operand(qk.inputs[0], &a.q_site, &a.q_sel, &qs, &qconst)) return false;
    if (a.q_site < 0) {
      if (qconst == nullptr) return fail("query operand has no recognisable source");
      a.q_const = true;
      a.q_val = an.get(qk.inputs[0]);  // value as it enters the product: query-side scalars are already folded in
    }
    if (!trace_operand(qk.inputs[1], &a.k_site, &a.k_sel, &ks, nullptr)) return false;
    if (a.k_site < 0) return fail("key operand does not come from a projection");
    a.scale *= (a.q_const ? 1.0 : qs) * ks;
    // softmax -> PV
    cur = sm.outputs[0];
    int pv_node = -1;
    for (int guard = 0; guard < 8 && pv_node < 0; ++guard) {
      auto c = consumers.find(cur);
      if (c == consumers.end() || c->second.empty()) return fail("Softmax output is unused");
      const OnnxNode& n = m.nodes[c->second[0]];
      if (n.op_type == "MatMul" && n.inputs[0] == cur) pv_node = c->second[0];
      else if (is_view_op(n.op_type)) cur = n.outputs[0];
      else return fail("unexpected operator '" + n.op_type + "' after Softmax");
    }
    if (pv_node < 0) return fail("no PV MatMul found after Softmax");
    if (!trace_operand(m.nodes[pv_node].inputs[1], &a.v_site, &a.v_sel, &vs, nullptr)) return false;
    if (a.v_site < 0) return fail("value operand does not come from a projection");
    if (fabs(vs - 1.0) > 1e-6) return fail("value operand is scaled");
    if (mask != nullptr) {
      if (!materialize(mask)) return fail("attention mask is not a foldable constant");
      const Shape& ms = mask->shape;
      if (ms.size() < 2 || ms[ms.size() - 1] != a.Tk || ms[ms.size() - 2] != a.Tq || numel(ms) != a.Tq * a.Tk)
        return fail("attention mask shape " + shape_str(ms) + " is not [Tq, Tk]");
      bool all_zero = true, causal = true;
      for (int64_t i = 0; i < a.Tq; ++i)
        for (int64_t j = 0; j < a.Tk; ++j) {
          const double v = mask->data[static_cast<size_t>(i * a.Tk + j)];
          if (v != 0.0) all_zero = false;
          if (j > i ? !(v < -1e4) : v != 0.0) causal = false;
        }
      if (!all_zero && !causal) return fail("attention mask is neither empty nor causal");
      a.mask = all_zero ? 0 : 1;
    }
    attns.push_back(a);
    return true;
  }

  bool tokenize() {
    std::unordered_map<int, int> attn_of_node, site_of_node;
    for (size_t i = 0; i < attns.size(); ++i) attn_of_node[attns[i].softmax_node] = static_cast<int>(i);
    for (size_t i = 0; i < sites.size(); ++i) site_of_node[sites[i].node] = static_cast<int>(i);
    bool has_argmax = false;
    for (const OnnxNode& n : m.nodes)
      if (n.op_type == "ArgMax" && !n.inputs.empty()) {
        AVal* v = an.get(n.inputs[0]);
        if (v != nullptr && v->tainted && !is_float_dt(v->dtype)) has_argmax = true;
      }
    for (size_t i = 0; i < m.nodes.size(); ++i) {
      const OnnxNode& n = m.nodes[i];
      if (n.outputs.empty()) continue;
      AVal* o = an.get(n.outputs[0]);
      if (o == nullptr || !o->tainted) continue;
      const int ni = static_cast<int>(i);
      Token t;
      t.node = ni;
      const std::string& op = n.op_type;
      if (site_of_node.count(ni)) {
        t.kind = sites[site_of_node[ni]].kind == 2 ? TK_CONV : TK_LIN;
        t.ref = site_of_node[ni];
      } else if (op == "LayerNormalization") {
        t.kind = TK_LN;
      } else if (op == "Softmax") {
        if (!attn_of_node.count(ni)) continue;
        t.kind = TK_SOFTMAX;
        t.ref = attn_of_node[ni];
      } else if (op == "Erf") {
        t.kind = TK_ACT; t.ref = 3;
      } else if (op == "Tanh") {
        t.kind = TK_ACT; t.ref = 2;
      } else if (op == "Gelu") {
        t.kind = TK_ACT;
        const OnnxAttr* a = n.attr("approximate");
        t.ref = (a != nullptr && a->s == "tanh") ? 2 : 3;
      } else if (op == "Sigmoid") {
        auto p = producer.find(n.inputs[0]);
        double c = 0.0;
        bool quick = false;
        if (p != producer.end() && m.nodes[p->second].op_type == "Mul") {
          const OnnxNode& mu = m.nodes[p->second];
          quick = (scalar_value(an.get(mu.inputs[1]), &c) || scalar_value(an.get(mu.inputs[0]), &c)) && fabs(c - 1.702) < 1e-3;
        }
        if (!quick) return fail("Sigmoid that is not part of QuickGELU (x * sigmoid(1.702 x))");
        t.kind = TK_ACT; t.ref = 1;
      } else if (op == "ReduceL2") {
        t.kind = TK_L2;
      } else if (op == "Gather") {
        AVal *d = an.get(n.inputs[0]), *ix = an.get(n.inputs[1]);
        if (!d->tainted && ix->tainted && d->shape.size() == 2 && d->init != nullptr) {
          t.kind = TK_EMBED;
          t.cval = d;
        } else if (d->tainted && is_float_dt(d->dtype)) {
          int64_t axis = n.attr_i("axis", 0);
          if (axis < 0) axis += static_cast<int64_t>(d->shape.size());
          if (ix->tainted) {
            if (!has_argmax) return fail("token select with an input-dependent index that is not an ArgMax");
            t.kind = TK_SEL; t.sel = 2;
          } else if (d->shape.size() == 3 && axis == 1 && numel(ix->shape) == 1) {
            std::vector<int64_t> iv;
            if (!an.ints_of(ix, &iv)) return fail("token select index is not a constant");
            const int64_t T = d->shape[1], g = iv[0] < 0 ? iv[0] + T : iv[0];
            if (T == 1) continue;
            if (g == 0) t.sel = 0;
            else if (g == T - 1) t.sel = 1;
            else return fail("token select at position " + std::to_string(g) + " (only first / last are supported)");
            t.kind = TK_SEL;
          } else continue;
        } else continue;
      } else if (op == "Slice") {
        AVal* d = an.get(n.inputs[0]);
        if (!(d->tainted && is_float_dt(d->dtype) && d->shape.size() == 3 && o->shape.size() == 3 && o->shape[1] == 1 &&
              d->shape[1] > 1 && o->shape[0] == d->shape[0] && o->shape[2] == d->shape[2]))
          continue;
        std::vector<int64_t> st;
        if (n.inputs.size() < 2 || !an.ints_of(an.get(n.inputs[1]), &st) || st.size() != 1) continue;
        const int64_t T = d->shape[1], g = st[0] < 0 ? st[0] + T : st[0];
        if (g == 0) t.sel = 0;
        else if (g == T - 1) t.sel = 1;
        else return fail("token slice at position " + std::to_string(g) + " (only first / last are supported)");
        t.kind = TK_SEL;
      } else if (op == "Concat") {
        int n_const = 0, n_taint = 0, const_pos = -1;
        for (size_t k = 0; k < n.inputs.size(); ++k) {
          AVal* v = an.get(n.inputs[k]);
          if (v == nullptr) continue;
          if (v->tainted) ++n_taint;
          else { ++n_const; const_pos = static_cast<int>(k); t.cval = v; }
        }
        if (!is_float_dt(o->dtype) || n_const == 0) continue;
        if (n_const != 1 || n_taint != 1 || const_pos != 0 || o->shape.size() != 3 || n.attr_i("axis", 0) != 1)
          return fail("unsupported Concat on the data path (only a prepended class token is recognised)");
        t.kind = TK_CLS;
      } else if (op == "Add") {
        if (bias_add_nodes.count(ni) || mask_add_nodes.count(ni)) continue;
        AVal *a = an.get(n.inputs[0]), *b = an.get(n.inputs[1]);
        AVal* c = !a->tainted ? a : (!b->tainted ? b : nullptr);
        if (c == nullptr || !is_float_dt(o->dtype)) continue;
        if (numel(c->shape) == 1) continue;  // scalar adds belong to activation formulas
        if (o->shape.size() == 3 && numel(c->shape) == o->shape[1] * o->shape[2]) {
          t.kind = TK_POS;
          t.cval = c;
        } else {
          return fail("unsupported constant Add on the data path (operand shape " + shape_str(c->shape) + ")");
        }
      } else {
        continue;
      }
      toks.push_back(t);
    }
    return true;
  }

  // ------------------------------------------------------------------------------------------------------
  // 3. grammar -> canonical tensors
  // ------------------------------------------------------------------------------------------------------
  OnnxTensor make_owned(const std::string& name, const Shape& dims, const std::vector<double>& d) {
    OnnxTensor t;
    t.name = name;
    t.dims = dims;
    t.data_type = 1;
    t.owned.resize(d.size() * 4);
    for (size_t i = 0; i < d.size(); ++i) {
      const float f = static_cast<float>(d[i]);
      memcpy(t.owned.data() + 4 * i, &f, 4);
    }
    t.nbytes = t.owned.size();
    return t;  // `data` is fixed up after the tensor has reached its final place in `out_tensors`
  }

  void add_tensor(OnnxTensor t, const std::string& source, bool transposed) {
    GraphBinding b;
    b.canonical = t.name;
    b.source = source;
    b.transposed = transposed;
    bindings.push_back(b);
    const std::string name = t.name;
    out_tensors[name] = std::move(t);
    OnnxTensor& ref = out_tensors[name];
    if (!ref.owned.empty()) ref.data = ref.owned.data();
  }

  // vector / table parameter: alias the initializer bytes when the value is one, otherwise store the folded value
  bool emit_plain(const std::string& name, AVal* v, const Shape& dims) {
    if (v == nullptr) return fail("missing tensor for '" + name + "'");
    if (numel(dims) != numel(v->shape)) return fail("'" + name + "': expected " + shape_str(dims) + ", graph has " + shape_str(v->shape));
    if (!v->has_data && v->init != nullptr && !v->init_t && is_float_dt(v->init->data_type) && v->init->owned.empty()) {
      OnnxTensor t;
      t.name = name;
      t.dims = dims;
      t.data_type = v->init->data_type;
      t.data = v->init->data;
      t.nbytes = v->init->nbytes;
      add_tensor(std::move(t), v->init_name.empty() ? "<constant>" : v->init_name, false);
      return true;
    }
    AVal tmp = *v;
    if (!materialize(&tmp)) return fail("'" + name + "' is not a foldable constant");
    add_tensor(make_owned(name, dims, tmp.data), v->init_name.empty() ? "<folded constant>" : v->init_name, false);
    return true;
  }

  // Constants that the exporter expanded along a FIXED batch axis (class token, attention-pool query of a graph
  // exported without dynamic_axes): every batch copy must be identical; emit the first one.
  bool emit_batch_constant(const std::string& name, AVal* v, const Shape& dims) {
    const int64_t want = numel(dims), b = an.batch0;
    if (v == nullptr || b <= 1 || numel(v->shape) != want * b) return emit_plain(name, v, dims);
    AVal tmp = *v;
    if (!materialize(&tmp)) return fail("'" + name + "' is not a foldable constant");
    for (int64_t r = 1; r < b; ++r)
      for (int64_t i = 0; i < want; ++i)
        if (tmp.data[static_cast<size_t>(r * want + i)] != tmp.data[static_cast<size_t>(i)])
          return fail("'" + name + "' differs between the batch copies of a fixed-batch export");
    tmp.data.resize(static_cast<size_t>(want));
    add_tensor(make_owned(name, dims, tmp.data), v->init_name.empty() ? "<folded constant>" : v->init_name, false);
    return true;
  }

  void emit_zeros(const std::string& name, int64_t n) {
    add_tensor(make_owned(name, {n}, std::vector<double>(static_cast<size_t>(n), 0.0)), "<zeros>", false);
  }

  // 2-D weight under a canonical [rows, cols] layout; `canon_nk`: canonical layout is [N(out), K(in)]
  bool emit_weight(const std::string& name, const LinearSite& s, bool canon_nk, const Shape& dims4 = {}) {
    const OnnxTensor* src = s.w->init;
    OnnxTensor t;
    t.name = name;
    t.dims = !dims4.empty() ? dims4 : (canon_nk ? Shape{s.N, s.K} : Shape{s.K, s.N});
    t.data_type = src->data_type;
    t.transposed = canon_nk ? s.w_kn : !s.w_kn;
    if (src->owned.empty()) {
      t.data = src->data;
      t.nbytes = src->nbytes;
    } else {
      t.owned = src->owned;
      t.nbytes = t.owned.size();
    }
    add_tensor(std::move(t), s.w->init_name, canon_nk ? s.w_kn : !s.w_kn);
    return true;
  }

  bool emit_bias(const std::string& name, const LinearSite& s, bool required) {
    if (s.bias == nullptr) {
      if (required) emit_zeros(name, s.N);
      return true;
    }
    return emit_plain(name, s.bias, {s.N});
  }

  bool emit_linear(const std::string& wname, const std::string& bname, const LinearSite& s, int64_t N, int64_t K) {
    if (s.N != N || s.K != K)
      return fail("'" + wname + "': expected a " + std::to_string(K) + "->" + std::to_string(N) + " projection, graph has " +
                  std::to_string(s.K) + "->" + std::to_string(s.N));
    return emit_weight(wname, s, true) && emit_bias(bname, s, true);
  }

  // several projections stacked along N (split q/k/v exports) -> one [sum N, K] fp32 tensor
  bool emit_stacked(const std::string& wname, const std::string& bname, const std::vector<int>& idx, int64_t K) {
    int64_t N = 0;
    std::string src;
    for (int si : idx) {
      if (sites[si].K != K) return fail("'" + wname + "': stacked projections disagree on the input width");
      N += sites[si].N;
      src += (src.empty() ? "concat(" : ", ") + sites[si].w->init_name;
    }
    src += ")";
    std::vector<double> w(static_cast<size_t>(N * K)), b(static_cast<size_t>(N), 0.0);
    int64_t row = 0;
    for (int si : idx) {
      const LinearSite& s = sites[si];
      std::vector<double> d;
      if (!read_tensor(*s.w->init, &d)) return fail("cannot read '" + s.w->init_name + "'");
      for (int64_t n = 0; n < s.N; ++n)
        for (int64_t k = 0; k < K; ++k)
          w[static_cast<size_t>((row + n) * K + k)] = s.w_kn ? d[static_cast<size_t>(k * s.N + n)] : d[static_cast<size_t>(n * K + k)];
      if (s.bias != nullptr) {
        AVal tmp = *s.bias;
        if (!materialize(&tmp)) return fail("bias of '" + s.w->init_name + "' is not a foldable constant");
        for (int64_t n = 0; n < s.N; ++n) b[static_cast<size_t>(row + n)] = tmp.data[static_cast<size_t>(n)];
      }
      row += s.N;
    }
    add_tensor(make_owned(wname, {N, K}, w), src, false);
    add_tensor(make_owned(bname, {N}, b), src + ".bias", false);
    return true;
  }

  struct Block {
    int ln1 = -1, ln2 = -1, attn = -1, proj = -1, fc1 = -1, fc2 = -1, act = 0;
    std::vector<int> qkv;  // 1 fused site or 3 sites in q,k,v order
  };

  bool emit_ln(const std::string& prefix, int tok_idx, int64_t D, double* eps) {
    const OnnxNode& n = m.nodes[toks[tok_idx].node];
    if (n.inputs.size() < 2) return fail("LayerNormalization without a scale");
    const int64_t axis = n.attr_i("axis", -1);
    AVal* x = an.get(n.inputs[0]);
    if (axis != -1 && axis != static_cast<int64_t>(x->shape.size()) - 1) return fail("LayerNormalization is not over the last axis");
    const double e = n.attr_f("epsilon", 1e-5f);
    if (*eps < 0) *eps = e;
    else if (fabs(*eps - e) > 1e-12) return fail("LayerNormalization layers use different epsilons");
    if (!emit_plain(prefix + ".weight", an.get(n.inputs[1]), {D})) return false;
    if (n.inputs.size() > 2 && !n.inputs[2].empty()) return emit_plain(prefix + ".bias", an.get(n.inputs[2]), {D});
    emit_zeros(prefix + ".bias", D);
    return true;
  }

  std::string tok_dump(size_t from) {
    std::string s;
    for (size_t i = from; i < toks.size() && i < from + 12; ++i) s += std::string(i > from ? " " : "") + tok_name(toks[i].kind);
    return s;
  }

  // parses `ln linear{1|3} softmax linear ln linear act linear` starting at toks[p]
  bool parse_block(size_t* p, int64_t D, Block* b) {
    size_t i = *p;
    auto expect = [&](TokKind k) {
      if (i < toks.size() && toks[i].kind == k) return true;
      fail(std::string("expected ") + tok_name(k) + " in a transformer block, found: " + tok_dump(i));
      return false;
    };
    if (!expect(TK_LN)) return false;
    b->ln1 = static_cast<int>(i++);
    std::vector<int> lin;
    while (i < toks.size() && toks[i].kind == TK_LIN) lin.push_back(toks[i++].ref);
    if (!expect(TK_SOFTMAX)) return false;
    b->attn = toks[i++].ref;
    const AttnSite& a = attns[b->attn];
    if (a.q_const) return fail("attention with a constant query inside the trunk");
    if (lin.size() == 1) {
      if (a.q_site != lin[0] || a.k_site != lin[0] || a.v_site != lin[0] || sites[lin[0]].N != 3 * D)
        return fail("fused projection does not feed q, k and v of its attention");
      if (a.q_sel != 0 || a.k_sel != 1 || a.v_sel != 2)
        return fail("fused qkv projection is not split in (q, k, v) order");
      b->qkv = lin;
    } else if (lin.size() == 3) {
      b->qkv = {a.q_site, a.k_site, a.v_site};
      std::set<int> have(lin.begin(), lin.end()), want(b->qkv.begin(), b->qkv.end());
      if (have != want || want.size() != 3) return fail("separate q/k/v projections do not match the attention operands");
      for (int si : b->qkv) if (sites[si].N != D) return fail("q/k/v projection width differs from the model width");
    } else {
      return fail("expected 1 or 3 projections in front of the attention, found " + std::to_string(lin.size()));
    }
    if (a.heads * a.hd != D) return fail("heads x head_dim does not equal the model width");
    if (!expect(TK_LIN)) return false;
    b->proj = toks[i++].ref;
    if (!expect(TK_LN)) return false;
    b->ln2 = static_cast<int>(i++);
    if (!expect(TK_LIN)) return false;
    b->fc1 = toks[i++].ref;
    if (!expect(TK_ACT)) return false;
    b->act = toks[i++].ref;
    if (!expect(TK_LIN)) return false;
    b->fc2 = toks[i++].ref;
    *p = i;
    return true;
  }

  bool is_block_start(size_t p) {
    if (p >= toks.size() || toks[p].kind != TK_LN) return false;
    size_t i = p + 1;
    while (i < toks.size() && toks[i].kind == TK_LIN) ++i;
    return i > p + 1 && i < toks.size() && toks[i].kind == TK_SOFTMAX && !attns[toks[i].ref].q_const;
  }

  struct Common {
    int64_t heads = 0, hd = 0, mlp = 0;
    int act = -1, mask = -1;
    double eps = -1;
  };

  bool emit_block(const std::string& p, bool timm, const Block& b, int64_t D, Common* c) {
    const AttnSite& a = attns[b.attn];
    if (c->heads == 0) { c->heads = a.heads; c->hd = a.hd; c->mask = a.mask; c->act = b.act; c->mlp = sites[b.fc1].N; }
    if (a.heads != c->heads || a.mask != c->mask || b.act != c->act || sites[b.fc1].N != c->mlp)
      return fail("transformer blocks differ in heads / mask / activation / MLP width");
    const double want = 1.0 / sqrt(static_cast<double>(a.hd));
    if (fabs(a.scale - want) > 1e-3 * want)
      return fail("attention scale " + std::to_string(a.scale) + " is not head_dim^-0.5 = " + std::to_string(want));
    if (!emit_ln(p + (timm ? ".norm1" : ".ln_1"), b.ln1, D, &c->eps)) return false;
    const std::string wq = p + (timm ? ".attn.qkv.weight" : ".attn.in_proj_weight");
    const std::string bq = p + (timm ? ".attn.qkv.bias" : ".attn.in_proj_bias");
    if (b.qkv.size() == 1) { if (!emit_linear(wq, bq, sites[b.qkv[0]], 3 * D, D)) return false; }
    else if (!emit_stacked(wq, bq, b.qkv, D)) return false;
    if (!emit_linear(p + (timm ? ".attn.proj.weight" : ".attn.out_proj.weight"),
                     p + (timm ? ".attn.proj.bias" : ".attn.out_proj.bias"), sites[b.proj], D, D)) return false;
    if (!emit_ln(p + (timm ? ".norm2" : ".ln_2"), b.ln2, D, &c->eps)) return false;
    if (!emit_linear(p + (timm ? ".mlp.fc1.weight" : ".mlp.c_fc.weight"), p + (timm ? ".mlp.fc1.bias" : ".mlp.c_fc.bias"),
                     sites[b.fc1], c->mlp, D)) return false;
    return emit_linear(p + (timm ? ".mlp.fc2.weight" : ".mlp.c_proj.weight"), p + (timm ? ".mlp.fc2.bias" : ".mlp.c_proj.bias"),
                       sites[b.fc2], D, c->mlp);
  }

  void common_meta(const Common& c, int64_t D, size_t layers) {
    char buf[64];
    snprintf(buf, sizeof buf, "%.9g", c.eps);
    out_meta["clipb200.eps"] = buf;
    out_meta["clipb200.heads"] = std::to_string(c.heads);
    out_meta["clipb200.act"] = std::to_string(c.act);
    out_meta["clipb200.width"] = std::to_string(D);
    out_meta["clipb200.layers"] = std::to_string(layers);
    out_meta["clipb200.mlp_dim"] = std::to_string(c.mlp);
    out_meta["clipb200.binding"] = "graph";
  }

  bool parse_vision() {
    size_t p = 0;
    if (toks.empty() || toks[0].kind != TK_CONV) return fail("vision graph does not start with a patch convolution: " + tok_dump(0));
    const LinearSite& conv = sites[toks[p++].ref];
    const OnnxNode& cn = m.nodes[conv.node];
    const Shape& ws = conv.w->shape;
    const int64_t D = ws[0], P = ws[2];
    if (D <= 0 || P <= 0) return fail("empty patch-embedding weight");
    std::vector<int64_t> strides = {1, 1}, pads = {0, 0, 0, 0};
    if (const OnnxAttr* a = cn.attr("strides")) strides = a->ints;
    if (const OnnxAttr* a = cn.attr("pads")) pads = a->ints;
    if (ws[1] != 3 || ws[2] != ws[3] || strides != std::vector<int64_t>{P, P} || pads != std::vector<int64_t>{0, 0, 0, 0} ||
        cn.attr_i("group", 1) != 1)
      return fail("first convolution is not a non-overlapping square patch embedding");
    AVal* cls = nullptr;
    if (p < toks.size() && toks[p].kind == TK_CLS) cls = toks[p++].cval;
    if (p >= toks.size() || toks[p].kind != TK_POS) return fail("no positional-embedding add after the patch embedding: " + tok_dump(p));
    AVal* pos = toks[p++].cval;
    const int64_t T = numel(pos->shape) / D;
    int ln_pre = -1;
    if (p < toks.size() && toks[p].kind == TK_LN && !is_block_start(p)) ln_pre = static_cast<int>(p++);
    std::vector<Block> blocks;
    while (is_block_start(p)) {
      Block b;
      if (!parse_block(&p, D, &b)) return false;
      blocks.push_back(b);
    }
    if (blocks.empty()) return fail("no transformer blocks recognised: " + tok_dump(p));
    // tail: `ln select` (open_clip default) or `select ln` (final_ln_after_pool / HF transformers) — the same function
    bool sel_before_ln = false;
    if (p + 1 < toks.size() && toks[p].kind == TK_SEL && toks[p].sel == 0 && toks[p + 1].kind == TK_LN) {
      sel_before_ln = true;
      ++p;
    }
    if (p >= toks.size() || toks[p].kind != TK_LN) return fail("no final LayerNorm after the blocks: " + tok_dump(p));
    const int ln_post = static_cast<int>(p++);
    bool map_tail = false;
    {
      size_t i = p;
      while (i < toks.size() && toks[i].kind == TK_LIN) ++i;
      map_tail = i > p && i < toks.size() && toks[i].kind == TK_SOFTMAX && attns[toks[i].ref].q_const;
    }
    Common c;
    const bool timm = map_tail;
    const std::string pre = timm ? "model.visual.trunk" : "model.visual";
    for (size_t i = 0; i < blocks.size(); ++i)
      if (!emit_block(pre + (timm ? ".blocks." : ".transformer.resblocks.") + std::to_string(i), timm, blocks[i], D, &c)) return false;
    if (c.mask != 0) return fail("vision tower with an attention mask");
    int64_t E = D;
    if (timm) {
      if (cls != nullptr || ln_pre >= 0 || sel_before_ln) return fail("attention-pool head combined with a class token / pre-norm is not supported");
      if (!emit_weight(pre + ".patch_embed.proj.weight", conv, true, ws) || !emit_bias(pre + ".patch_embed.proj.bias", conv, true)) return false;
      if (!emit_plain(pre + ".pos_embed", pos, {1, T, D})) return false;
      if (!emit_ln(pre + ".norm", ln_post, D, &c.eps)) return false;
      std::vector<int> kv;
      while (p < toks.size() && toks[p].kind == TK_LIN) kv.push_back(toks[p++].ref);
      const AttnSite& a = attns[toks[p++].ref];
      if (a.Tq != 1 || a.heads != c.heads || a.hd != c.hd || a.mask != 0) return fail("attention-pool head has unexpected geometry");
      const std::string ap = pre + ".attn_pool";
      if (kv.size() == 1) {
        if (a.k_site != kv[0] || a.v_site != kv[0] || a.k_sel != 0 || a.v_sel != 1 || sites[kv[0]].N != 2 * D)
          return fail("attention-pool kv projection is not split in (k, v) order");
        if (!emit_linear(ap + ".kv.weight", ap + ".kv.bias", sites[kv[0]], 2 * D, D)) return false;
      } else if (kv.size() == 2) {
        if (!emit_stacked(ap + ".kv.weight", ap + ".kv.bias", {a.k_site, a.v_site}, D)) return false;
      } else return fail("attention-pool head: expected 1 or 2 key/value projections");
      AVal q = *a.q_val;
      if (!materialize(&q) || (numel(q.shape) != D && numel(q.shape) != D * an.batch0))
        return fail("attention-pool query is not a foldable [1, H, 1, hd] constant");
      if (numel(q.shape) != D) {  // fixed-batch export: identical copies along the batch axis
        for (int64_t i = D; i < numel(q.shape); ++i)
          if (q.data[static_cast<size_t>(i)] != q.data[static_cast<size_t>(i % D)]) return fail("attention-pool query differs between batch copies");
        q.data.resize(static_cast<size_t>(D));
      }
      // a.scale holds the scalars on the key side and behind QK^T; the engine applies none, so fold all of it into q
      std::vector<double> qd(q.data);
      const double want = 1.0 / sqrt(static_cast<double>(a.hd));
      // the query-side scalars are already inside q_val: the total scale is only checkable up to them, so require the
      // symmetric sqrt(scale) split or a k-side/post scale of exactly hd^-0.5 or 1
      const double rest = a.scale;
      const bool ok = fabs(rest - want) < 1e-3 * want || fabs(rest - sqrt(want)) < 1e-3 * sqrt(want) || fabs(rest - 1.0) < 1e-6;
      if (!ok) return fail("attention-pool scale " + std::to_string(rest) + " is not recognised");
      for (double& d : qd) d *= rest;
      add_tensor(make_owned("clipb200.map_query", {D}, qd), "<folded constant>", false);
      auto expect = [&](TokKind k) {
        if (p < toks.size() && toks[p].kind == k) return true;
        fail(std::string("expected ") + tok_name(k) + " in the attention-pool head, found: " + tok_dump(p));
        return false;
      };
      if (!expect(TK_LIN)) return false;
      if (!emit_linear(ap + ".proj.weight", ap + ".proj.bias", sites[toks[p++].ref], D, D)) return false;
      if (!expect(TK_LN)) return false;
      if (!emit_ln(ap + ".norm", static_cast<int>(p++), D, &c.eps)) return false;
      if (!expect(TK_LIN)) return false;
      const LinearSite& f1 = sites[toks[p++].ref];
      if (!emit_linear(ap + ".mlp.fc1.weight", ap + ".mlp.fc1.bias", f1, c.mlp, D)) return false;
      if (!expect(TK_ACT)) return false;
      if (toks[p++].ref != c.act) return fail("attention-pool MLP uses a different activation");
      if (!expect(TK_LIN)) return false;
      if (!emit_linear(ap + ".mlp.fc2.weight", ap + ".mlp.fc2.bias", sites[toks[p++].ref], D, c.mlp)) return false;
      out_meta["clipb200.pool"] = "map";
      out_meta["clipb200.family"] = "timm";
    } else {
      if (cls == nullptr || ln_pre < 0) return fail("projection-head vision tower without class token / pre-norm is not supported");
      if (conv.bias != nullptr) return fail("class-token vision tower with a patch-embedding bias is not supported");
      if (!sel_before_ln) {
        if (p >= toks.size() || toks[p].kind != TK_SEL || toks[p].sel != 0) return fail("expected class-token pooling after the final LayerNorm: " + tok_dump(p));
        ++p;
      }
      if (p >= toks.size() || toks[p].kind != TK_LIN) return fail("expected the output projection: " + tok_dump(p));
      const LinearSite& head = sites[toks[p++].ref];
      if (head.K != D || head.bias != nullptr) return fail("output projection must be a bias-free D->E matrix");
      E = head.N;
      if (!emit_weight(pre + ".conv1.weight", conv, true, ws)) return false;
      if (!emit_batch_constant(pre + ".class_embedding", cls, {D})) return false;
      if (T * D != numel(pos->shape)) return fail("positional embedding size");
      if (!emit_plain(pre + ".positional_embedding", pos, {T, D})) return false;
      if (!emit_ln(pre + ".ln_pre", ln_pre, D, &c.eps)) return false;
      if (!emit_ln(pre + ".ln_post", ln_post, D, &c.eps)) return false;
      if (!emit_weight(pre + ".proj", head, false)) return false;
      out_meta["clipb200.pool"] = "cls";
      out_meta["clipb200.family"] = "clip";
    }
    if (p >= toks.size() || toks[p].kind != TK_L2) return fail("graph output is not L2-normalised: " + tok_dump(p));
    common_meta(c, D, blocks.size());
    out_meta["clipb200.tower"] = "vision";
    out_meta["clipb200.patch"] = std::to_string(P);
    out_meta["clipb200.embed_dim"] = std::to_string(E);
    return true;
  }

  bool parse_text() {
    size_t p = 0;
    if (toks.empty() || toks[0].kind != TK_EMBED) return fail("text graph does not start with a token-embedding gather: " + tok_dump(0));
    AVal* table = toks[p++].cval;
    const int64_t V = table->shape[0], D = table->shape[1];
    if (V <= 0 || D <= 0) return fail("empty token-embedding table");
    if (p >= toks.size() || toks[p].kind != TK_POS) return fail("no positional-embedding add after the token embedding: " + tok_dump(p));
    AVal* pos = toks[p++].cval;
    const int64_t T = numel(pos->shape) / D;
    std::vector<Block> blocks;
    while (is_block_start(p)) {
      Block b;
      if (!parse_block(&p, D, &b)) return false;
      blocks.push_back(b);
    }
    if (blocks.empty()) return fail("no transformer blocks recognised: " + tok_dump(p));
    if (p >= toks.size() || toks[p].kind != TK_LN) return fail("no final LayerNorm after the blocks: " + tok_dump(p));
    const int ln_final = static_cast<int>(p++);
    if (p >= toks.size() || toks[p].kind != TK_SEL) return fail("no token pooling after the final LayerNorm: " + tok_dump(p));
    const int sel = toks[p++].sel;
    if (sel == 0) return fail("first-token pooling in a text tower is not supported");
    if (p >= toks.size() || toks[p].kind != TK_LIN) return fail("expected the text projection: " + tok_dump(p));
    const LinearSite& head = sites[toks[p++].ref];
    if (head.K != D) return fail("text projection input width differs from the model width");
    if (p >= toks.size() || toks[p].kind != TK_L2) return fail("graph output is not L2-normalised: " + tok_dump(p));
    Common c;
    const std::string pre = "model";
    for (size_t i = 0; i < blocks.size(); ++i)
      if (!emit_block(pre + ".transformer.resblocks." + std::to_string(i), false, blocks[i], D, &c)) return false;
    if (!emit_plain(pre + ".token_embedding.weight", table, {V, D})) return false;
    if (!emit_plain(pre + ".positional_embedding", pos, {T, D})) return false;
    if (!emit_ln(pre + ".ln_final", ln_final, D, &c.eps)) return false;
    if (head.bias != nullptr) {
      if (!emit_weight(pre + ".text_projection.weight", head, true) || !emit_bias(pre + ".text_projection.bias", head, true)) return false;
    } else if (!emit_weight(pre + ".text_projection", head, false)) return false;
    common_meta(c, D, blocks.size());
    out_meta["clipb200.tower"] = "text";
    out_meta["clipb200.family"] = "clip";
    out_meta["clipb200.causal"] = c.mask == 1 ? "1" : "0";
    out_meta["clipb200.pool"] = sel == 2 ? "argmax" : "last";
    out_meta["clipb200.context_length"] = std::to_string(T);
    out_meta["clipb200.vocab_size"] = std::to_string(V);
    out_meta["clipb200.embed_dim"] = std::to_string(head.N);
    return true;
  }

  bool run() {
    if (m.input_infos.size() != 1) return fail("expected exactly one graph input, found " + std::to_string(m.input_infos.size()));
    if (!an.run()) return fail(an.err);
    index_graph();
    if (!find_sites()) return false;
    for (size_t i = 0; i < m.nodes.size(); ++i)
      if (m.nodes[i].op_type == "Softmax" && !analyse_softmax(static_cast<int>(i))) {
        err += " (Softmax node " + std::to_string(i) + ")";
        return false;
      }
    if (!tokenize()) return false;
    const bool text = !is_float_dt(m.input_infos[0].elem_type);
    return text ? parse_text() : parse_vision();
  }
};

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// Re-parameterised FastViT trunks (MobileCLIP2): the conv graph is not parsed by the recogniser above, and it does not
// have to be: `torch.onnx.export` keeps every Conv / BatchNorm / bias / layer-scale / head parameter under its module
// name, so those bind by name.  What it does NOT keep are the two `nn.Linear` weights of every attention block: a Linear
// applied to [B, N, C] tokens becomes `MatMul(x, W^T)` with the pre-transposed constant renamed `onnx::MatMul_<n>`.
// They are found from the named tensors next to them, following the graph's edges:
//   qkv :  BatchNormalization(scale = "<block>.norm.weight") -> {Reshape | Flatten | Transpose | ...} -> MatMul(B = const [C, 3C])
//   proj:  Add(_, "<block>.token_mixer.proj.bias") <- MatMul(B = const [C, C])
// and registered under timm's names as transposed aliases of the same bytes.
bool bind_fastvit_graph(OnnxModel* m, std::string* err) {
  if (m->nodes.empty()) return true;   // initializer-only file: every name is already there
  std::map<std::string, std::vector<int>> consumers;
  std::map<std::string, int> producer;
  for (size_t i = 0; i < m->nodes.size(); ++i) {
    for (const std::string& in : m->nodes[i].inputs) consumers[in].push_back(static_cast<int>(i));
    for (const std::string& out : m->nodes[i].outputs) producer[out] = static_cast<int>(i);
  }
  auto const_b = [&](const OnnxNode& n, int64_t rows, int64_t cols) -> const OnnxTensor* {
    if (n.op_type != "MatMul" || n.inputs.size() != 2) return nullptr;
    const OnnxTensor* t = m->find(n.inputs[1]);
    if (t == nullptr || t->dims.size() != 2 || t->dims[0] != rows || t->dims[1] != cols || !is_float_dt(t->data_type)) return nullptr;
    return t;
  };
  auto alias = [&](const std::string& name, const OnnxTensor& src, int64_t rows, int64_t cols) {
    OnnxTensor t;
    t.name = name;
    t.dims = {rows, cols};
    t.data_type = src.data_type;
    t.data = src.data;
    t.nbytes = src.nbytes;
    t.owned = src.owned;
    if (!t.owned.empty()) t.data = nullptr;
    t.transposed = true;   // bytes are [cols, rows]
    m->initializers[name] = std::move(t);
    OnnxTensor& ref = m->initializers[name];
    if (!ref.owned.empty()) ref.data = ref.owned.data();
  };
  int bound = 0;
  for (const OnnxNode& bn : m->nodes) {
    if (bn.op_type != "BatchNormalization" || bn.inputs.size() < 5 || bn.outputs.empty()) continue;
    const std::string& scale = bn.inputs[1];
    const std::string suffix = ".norm.weight";
    if (scale.size() <= suffix.size() || scale.compare(scale.size() - suffix.size(), suffix.size(), suffix) != 0) continue;
    const std::string block = scale.substr(0, scale.size() - suffix.size());
    if (m->has(block + ".token_mixer.qkv.weight")) continue;   // the exporter kept the names (or an earlier pass bound them)
    const OnnxTensor* sc = m->find(scale);
    if (sc == nullptr || sc->dims.size() != 1) continue;
    const int64_t C = sc->dims[0];
    // qkv: breadth-first from the BatchNorm output through shape-only operators
    const OnnxTensor* wq = nullptr;
    std::vector<std::string> frontier = {bn.outputs[0]};
    for (int depth = 0; depth < 6 && wq == nullptr && !frontier.empty(); ++depth) {
      std::vector<std::string> next;
      for (const std::string& tname : frontier) {
        for (int ci : consumers[tname]) {
          const OnnxNode& n = m->nodes[static_cast<size_t>(ci)];
          if (n.inputs.empty() || n.outputs.empty() || n.inputs[0] != tname) continue;   // only along the data input
          if ((wq = const_b(n, C, 3 * C)) != nullptr) break;
          if (n.op_type == "Reshape" || n.op_type == "Flatten" || n.op_type == "Transpose" || n.op_type == "Identity" ||
              n.op_type == "Squeeze" || n.op_type == "Unsqueeze")
            next.push_back(n.outputs[0]);
        }
        if (wq != nullptr) break;
      }
      frontier.swap(next);
    }
    if (wq == nullptr) {
      *err = "FastViT attention block '" + block + "': no MatMul with a constant [C, 3C] operand follows its BatchNormalization";
      return false;
    }
    // proj: the MatMul feeding the Add that carries the named bias
    const OnnxTensor* wp = nullptr;
    const std::string pbias = block + ".token_mixer.proj.bias";
    for (int ci : consumers[pbias]) {
      const OnnxNode& add = m->nodes[static_cast<size_t>(ci)];
      if (add.op_type != "Add" || add.inputs.size() != 2) continue;
      const std::string& other = add.inputs[0] == pbias ? add.inputs[1] : add.inputs[0];
      auto it = producer.find(other);
      if (it != producer.end()) wp = const_b(m->nodes[static_cast<size_t>(it->second)], C, C);
      if (wp != nullptr) break;
    }
    if (wp == nullptr) {
      *err = "FastViT attention block '" + block + "': no MatMul with a constant [C, C] operand feeds the Add of '" + pbias + "'";
      return false;
    }
    alias(block + ".token_mixer.qkv.weight", *wq, 3 * C, C);
    alias(block + ".token_mixer.proj.weight", *wp, C, C);
    ++bound;
  }
  if (bound > 0) m->metadata["clipb200.fastvit_graph_linears"] = std::to_string(2 * bound);
  return true;
}

bool graph_needs_recognition(const OnnxModel& m) {
  if (!m.meta("clipb200.family").empty()) return false;
  for (const OnnxNode& n : m.nodes)
    if (n.op_type == "MatMul" || n.op_type == "Gemm" || n.op_type == "Conv") return true;
  return false;
}

bool tensor_to_f32(const OnnxTensor& t, std::vector<float>* out) {
  if (!is_float_dt(t.data_type)) return false;
  std::vector<double> d;
  if (!read_tensor(t, &d)) return false;
  out->resize(d.size());
  if (t.transposed && t.dims.size() == 2) {
    const int64_t r = t.dims[0], c = t.dims[1];  // canonical [r, c]; bytes are [c, r]
    for (int64_t i = 0; i < r; ++i)
      for (int64_t j = 0; j < c; ++j) (*out)[static_cast<size_t>(i * c + j)] = static_cast<float>(d[static_cast<size_t>(j * r + i)]);
  } else {
    for (size_t i = 0; i < d.size(); ++i) (*out)[i] = static_cast<float>(d[i]);
  }
  return true;
}

bool recognize_graph(OnnxModel* m, std::string* err, std::vector<GraphBinding>* bindings_or_null) {
  Recognizer r(*m);
  if (!r.run()) {
    *err = "graph recogniser: " + (r.err.empty() ? std::string("unknown failure") : r.err);
    return false;
  }
  for (auto& kv : r.out_tensors) {
    OnnxTensor t = std::move(kv.second);
    const bool own = !t.owned.empty();
    auto it = m->initializers.find(kv.first);
    if (it != m->initializers.end()) m->initializers.erase(it);
    auto ins = m->initializers.emplace(kv.first, std::move(t));
    if (own) ins.first->second.data = ins.first->second.owned.data();
  }
  for (const auto& kv : r.out_meta) m->metadata[kv.first] = kv.second;
  if (bindings_or_null != nullptr) *bindings_or_null = r.bindings;
  return true;
}

}  // namespace clipb200

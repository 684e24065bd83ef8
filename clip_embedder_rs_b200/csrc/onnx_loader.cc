#include "onnx_loader.h"

#include <errno.h>
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace clipb200 {

MappedFile::~MappedFile() {
  if (base != nullptr && size > 0) munmap(base, size);
}

static std::shared_ptr<MappedFile> map_file(const std::string& path, std::string* err) {
  int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) {
    *err = "cannot open '" + path + "': " + strerror(errno);
    return nullptr;
  }
  struct stat st;
  if (fstat(fd, &st) != 0) {
    *err = "cannot stat '" + path + "'";
    close(fd);
    return nullptr;
  }
  auto mf = std::make_shared<MappedFile>();
  mf->size = static_cast<size_t>(st.st_size);
  if (mf->size > 0) {
    void* p = mmap(nullptr, mf->size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (p == MAP_FAILED) {
      *err = "cannot mmap '" + path + "'";
      close(fd);
      return nullptr;
    }
    mf->base = p;
  }
  close(fd);
  return mf;
}

namespace {

struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  bool done() const { return p >= end || !ok; }
  uint64_t varint() {
    uint64_t r = 0;
    int shift = 0;
    while (p < end && shift < 70) {
      const uint8_t b = *p++;
      r |= static_cast<uint64_t>(b & 0x7F) << shift;
      if (!(b & 0x80)) return r;
      shift += 7;
    }
    ok = false;
    return 0;
  }
  // reads a key; returns field number, sets wire type
  int key(int* wire) {
    const uint64_t k = varint();
    *wire = static_cast<int>(k & 7);
    return static_cast<int>(k >> 3);
  }
  Cursor sub() {
    const uint64_t n = varint();
    if (!ok || n > static_cast<uint64_t>(end - p)) {
      ok = false;
      return Cursor{p, p, false};
    }
    Cursor c{p, p + n, true};
    p += n;
    return c;
  }
  void skip(int wire) {
    switch (wire) {
      case 0: varint(); break;
      case 1: if (end - p >= 8) p += 8; else ok = false; break;
      case 2: sub(); break;
      case 5: if (end - p >= 4) p += 4; else ok = false; break;
      default: ok = false;
    }
  }
  std::string str() {
    Cursor c = sub();
    return std::string(reinterpret_cast<const char*>(c.p), static_cast<size_t>(c.end - c.p));
  }
};

void parse_kv(Cursor c, std::string* k, std::string* v) {
  while (!c.done()) {
    int w;
    const int f = c.key(&w);
    if (f == 1 && w == 2) *k = c.str();
    else if (f == 2 && w == 2) *v = c.str();
    else c.skip(w);
  }
}

struct ExtRef {
  std::string location;
  int64_t offset = 0;
  int64_t length = -1;
  bool present = false;
};

bool parse_tensor(Cursor c, OnnxTensor* t, ExtRef* ext) {
  std::vector<float> float_data;
  std::vector<int64_t> int64_data;
  std::vector<int32_t> int32_data;
  while (!c.done()) {
    int w;
    const int f = c.key(&w);
    if (f == 1) {  // dims
      if (w == 2) {
        Cursor d = c.sub();
        while (!d.done()) t->dims.push_back(static_cast<int64_t>(d.varint()));
      } else {
        t->dims.push_back(static_cast<int64_t>(c.varint()));
      }
    } else if (f == 2 && w == 0) {
      t->data_type = static_cast<int>(c.varint());
    } else if (f == 8 && w == 2) {
      t->name = c.str();
    } else if (f == 9 && w == 2) {  // raw_data
      Cursor r = c.sub();
      t->data = r.p;
      t->nbytes = static_cast<size_t>(r.end - r.p);
    } else if (f == 13 && w == 2) {  // external_data entry
      std::string k, v;
      parse_kv(c.sub(), &k, &v);
      ext->present = true;
      if (k == "location") ext->location = v;
      else if (k == "offset") ext->offset = atoll(v.c_str());
      else if (k == "length") ext->length = atoll(v.c_str());
    } else if (f == 4) {  // float_data
      if (w == 2) {
        Cursor r = c.sub();
        const size_t n = static_cast<size_t>(r.end - r.p) / 4;
        const size_t old = float_data.size();
        float_data.resize(old + n);
        if (n > 0) memcpy(float_data.data() + old, r.p, n * 4);
      } else if (w == 5) {
        if (c.end - c.p < 4) { c.ok = false; break; }
        float v;
        memcpy(&v, c.p, 4);
        c.p += 4;
        float_data.push_back(v);
      } else {
        c.skip(w);
      }
    } else if (f == 5) {  // int32_data (also carries bool / int8 / uint8 / fp16 bit patterns)
      if (w == 2) {
        Cursor r = c.sub();
        while (!r.done()) int32_data.push_back(static_cast<int32_t>(r.varint()));
      } else {
        int32_data.push_back(static_cast<int32_t>(c.varint()));
      }
    } else if (f == 7) {  // int64_data
      if (w == 2) {
        Cursor r = c.sub();
        while (!r.done()) int64_data.push_back(static_cast<int64_t>(r.varint()));
      } else {
        int64_data.push_back(static_cast<int64_t>(c.varint()));
      }
    } else {
      c.skip(w);
    }
  }
  if (t->data == nullptr && !ext->present) {
    if (!float_data.empty()) {
      t->owned.resize(float_data.size() * 4);
      memcpy(t->owned.data(), float_data.data(), t->owned.size());
    } else if (!int64_data.empty()) {
      t->owned.resize(int64_data.size() * 8);
      memcpy(t->owned.data(), int64_data.data(), t->owned.size());
    } else if (!int32_data.empty()) {
      if (t->data_type == 6) {
        t->owned.resize(int32_data.size() * 4);
        memcpy(t->owned.data(), int32_data.data(), t->owned.size());
      } else if (t->data_type == 9 || t->data_type == 2 || t->data_type == 3) {
        t->owned.resize(int32_data.size());
        for (size_t i = 0; i < int32_data.size(); ++i) t->owned[i] = static_cast<uint8_t>(int32_data[i]);
      } else if (t->data_type == 10 || t->data_type == 16) {
        t->owned.resize(int32_data.size() * 2);
        for (size_t i = 0; i < int32_data.size(); ++i) {
          const uint16_t h = static_cast<uint16_t>(int32_data[i]);
          memcpy(t->owned.data() + 2 * i, &h, 2);
        }
      }
    }
    t->data = t->owned.data();
    t->nbytes = t->owned.size();
  }
  return c.ok;
}

// ValueInfoProto{1 name, 2 type{1 tensor_type{1 elem_type, 2 shape{1 dim{1 dim_value | 2 dim_param}}}}}
void parse_value_info(Cursor c, OnnxValueInfo* vi) {
  while (!c.done()) {
    int w;
    const int f = c.key(&w);
    if (f == 1 && w == 2) {
      vi->name = c.str();
    } else if (f == 2 && w == 2) {
      Cursor ty = c.sub();
      while (!ty.done()) {
        int tw;
        const int tf = ty.key(&tw);
        if (tf == 1 && tw == 2) {
          Cursor tt = ty.sub();
          while (!tt.done()) {
            int ew;
            const int ef = tt.key(&ew);
            if (ef == 1 && ew == 0) {
              vi->elem_type = static_cast<int>(tt.varint());
            } else if (ef == 2 && ew == 2) {
              Cursor sh = tt.sub();
              while (!sh.done()) {
                int sw;
                const int sf = sh.key(&sw);
                if (sf == 1 && sw == 2) {
                  Cursor d = sh.sub();
                  int64_t value = -1;
                  while (!d.done()) {
                    int dw;
                    const int df = d.key(&dw);
                    if (df == 1 && dw == 0) value = static_cast<int64_t>(d.varint());
                    else d.skip(dw);
                  }
                  vi->dims.push_back(value);
                } else {
                  sh.skip(sw);
                }
              }
            } else {
              tt.skip(ew);
            }
          }
        } else {
          ty.skip(tw);
        }
      }
    } else {
      c.skip(w);
    }
  }
}

bool parse_tensor(Cursor c, OnnxTensor* t, ExtRef* ext);

void parse_attr(Cursor c, std::string* name, OnnxAttr* a) {
  while (!c.done()) {
    int w;
    const int f = c.key(&w);
    if (f == 1 && w == 2) *name = c.str();
    else if (f == 2 && w == 5) {
      if (c.end - c.p < 4) { c.ok = false; break; }
      memcpy(&a->f, c.p, 4);
      c.p += 4;
      if (a->type == 0) a->type = 1;
    }
    else if (f == 3 && w == 0) { a->i = static_cast<int64_t>(c.varint()); if (a->type == 0) a->type = 2; }
    else if (f == 4 && w == 2) { a->s = c.str(); if (a->type == 0) a->type = 3; }
    else if (f == 5 && w == 2) {
      a->t = std::make_shared<OnnxTensor>();
      ExtRef ext;
      if (!parse_tensor(c.sub(), a->t.get(), &ext) || ext.present) a->t.reset();  // constants are always inline
      if (a->type == 0) a->type = 4;
    } else if (f == 7) {
      if (w == 2) {
        Cursor r = c.sub();
        while (r.end - r.p >= 4) { float v; memcpy(&v, r.p, 4); r.p += 4; a->floats.push_back(v); }
      } else if (w == 5) {
        if (c.end - c.p < 4) { c.ok = false; break; }
        float v;
        memcpy(&v, c.p, 4);
        c.p += 4;
        a->floats.push_back(v);
      } else c.skip(w);
    } else if (f == 8) {
      if (w == 2) {
        Cursor r = c.sub();
        while (!r.done()) a->ints.push_back(static_cast<int64_t>(r.varint()));
      } else if (w == 0) a->ints.push_back(static_cast<int64_t>(c.varint()));
      else c.skip(w);
    } else if (f == 20 && w == 0) {
      a->type = static_cast<int>(c.varint());
    } else {
      c.skip(w);
    }
  }
}

void parse_node(Cursor c, OnnxNode* n) {
  while (!c.done()) {
    int w;
    const int f = c.key(&w);
    if (f == 1 && w == 2) n->inputs.push_back(c.str());
    else if (f == 2 && w == 2) n->outputs.push_back(c.str());
    else if (f == 3 && w == 2) n->name = c.str();
    else if (f == 4 && w == 2) n->op_type = c.str();
    else if (f == 5 && w == 2) {
      std::string name;
      OnnxAttr a;
      parse_attr(c.sub(), &name, &a);
      n->attrs.emplace(std::move(name), std::move(a));
    } else c.skip(w);
  }
}

size_t dtype_size(int dt) {
  switch (dt) {
    case 1: return 4;   // float
    case 6: return 4;   // int32
    case 7: return 8;   // int64
    case 10: return 2;  // float16
    case 16: return 2;  // bfloat16
    case 11: return 8;  // double
    case 9: return 1;   // bool
    default: return 0;
  }
}

}  // namespace

const OnnxTensor* OnnxModel::find(const std::string& name) const {
  auto it = initializers.find(name);
  return it == initializers.end() ? nullptr : &it->second;
}

std::string OnnxModel::meta(const std::string& key, const std::string& dflt) const {
  auto it = metadata.find(key);
  return it == metadata.end() ? dflt : it->second;
}

// A relative path whose components never step out of the directory it is joined to.
static bool external_location_is_safe(const std::string& loc) {
  if (loc.empty() || loc[0] == '/' || loc[0] == '\\' || loc.find('\0') != std::string::npos) return false;
  if (loc.size() >= 2 && loc[1] == ':') return false;  // drive letter
  size_t i = 0;
  int depth = 0;
  while (i <= loc.size()) {
    size_t j = loc.find_first_of("/\\", i);
    if (j == std::string::npos) j = loc.size();
    const std::string part = loc.substr(i, j - i);
    if (part == "..") {
      if (--depth < 0) return false;
    } else if (!part.empty() && part != ".") {
      ++depth;
    }
    i = j + 1;
  }
  return depth > 0;
}

bool load_onnx(const std::string& path, OnnxModel* model, std::string* err) {
  auto mf = map_file(path, err);
  if (!mf) return false;
  if (mf->size == 0) {
    *err = "'" + path + "' is empty";
    return false;
  }
  model->files.push_back(mf);
  std::string dir = ".";
  const size_t slash = path.find_last_of('/');
  if (slash != std::string::npos) dir = path.substr(0, slash);
  std::map<std::string, std::shared_ptr<MappedFile>> ext_files;
  std::vector<OnnxValueInfo> graph_inputs;
  bool saw_graph = false;

  Cursor c{static_cast<const uint8_t*>(mf->base), static_cast<const uint8_t*>(mf->base) + mf->size, true};
  while (!c.done()) {
    int w;
    const int f = c.key(&w);
    if (f == 7 && w == 2) {  // graph
      saw_graph = true;
      Cursor g = c.sub();
      while (!g.done()) {
        int gw;
        const int gf = g.key(&gw);
        if (gf == 5 && gw == 2) {  // initializer
          OnnxTensor t;
          ExtRef ext;
          if (!parse_tensor(g.sub(), &t, &ext)) {
            *err = "malformed TensorProto in '" + path + "'";
            return false;
          }
          bool dims_ok = t.dims.size() <= 8;
          {
            int64_t prod = 1;
            for (int64_t d : t.dims) {
              if (d < 0 || d > (int64_t(1) << 40) || (d > 0 && prod > (int64_t(1) << 46) / d)) { dims_ok = false; break; }
              prod *= d;
            }
          }
          if (!dims_ok) {
            *err = "initializer '" + t.name + "' has invalid dimensions";
            return false;
          }
          const size_t esz = dtype_size(t.data_type);
          if (esz == 0) {
            *err = "initializer '" + t.name + "' has unsupported data_type " + std::to_string(t.data_type);
            return false;
          }
          const size_t want = static_cast<size_t>(t.numel()) * esz;
          if (ext.present) {
            if (ext.location.empty()) {
              *err = "initializer '" + t.name + "' has external data without a location";
              return false;
            }
            // onnxruntime refuses external-data locations that leave the model directory; so does this loader (an
            // untrusted .onnx must not be able to map /etc/... or ../secrets through `location`)
            if (!external_location_is_safe(ext.location)) {
              *err = "initializer '" + t.name + "' external data location '" + ext.location +
                     "' is absolute or escapes the model directory";
              return false;
            }
            auto it = ext_files.find(ext.location);
            if (it == ext_files.end()) {
              auto ef = map_file(dir + "/" + ext.location, err);
              if (!ef) return false;
              model->files.push_back(ef);
              it = ext_files.emplace(ext.location, ef).first;
            }
            const size_t len = ext.length >= 0 ? static_cast<size_t>(ext.length) : want;
            if (ext.offset < 0 || static_cast<size_t>(ext.offset) > it->second->size ||
                len > it->second->size - static_cast<size_t>(ext.offset) || len < want) {
              *err = "initializer '" + t.name + "' external data range is outside '" + ext.location + "'";
              return false;
            }
            t.data = static_cast<const uint8_t*>(it->second->base) + ext.offset;
            t.nbytes = len;
          } else if (t.nbytes < want) {
            *err = "initializer '" + t.name + "' holds " + std::to_string(t.nbytes) + " bytes, dims need " +
                   std::to_string(want);
            return false;
          }
          model->initializers.emplace(t.name, std::move(t));
        } else if ((gf == 11 || gf == 12) && gw == 2) {
          OnnxValueInfo vi;
          parse_value_info(g.sub(), &vi);
          if (gf == 11) graph_inputs.push_back(std::move(vi));
          else model->outputs.push_back(vi.name);
        } else if (gf == 1 && gw == 2) {
          OnnxNode n;
          parse_node(g.sub(), &n);
          model->nodes.push_back(std::move(n));
        } else {
          g.skip(gw);
        }
      }
      if (!g.ok) {
        *err = "malformed GraphProto in '" + path + "'";
        return false;
      }
    } else if (f == 8 && w == 2) {  // opset_import
      Cursor o = c.sub();
      std::string domain;
      int64_t version = 0;
      while (!o.done()) {
        int ow;
        const int of = o.key(&ow);
        if (of == 1 && ow == 2) domain = o.str();
        else if (of == 2 && ow == 0) version = static_cast<int64_t>(o.varint());
        else o.skip(ow);
      }
      if (domain.empty()) model->opset = version;
    } else if (f == 14 && w == 2) {  // metadata_props
      std::string k, v;
      parse_kv(c.sub(), &k, &v);
      model->metadata[k] = v;
    } else {
      c.skip(w);
    }
  }
  if (!c.ok || !saw_graph) {
    *err = "'" + path + "' is not a valid ONNX ModelProto";
    return false;
  }
  // Exporters de-duplicate identical initializers through Identity nodes: alias them.
  for (const OnnxNode& n : model->nodes) {
    if (n.op_type == "Identity" && n.inputs.size() == 1 && n.outputs.size() == 1) {
      auto it = model->initializers.find(n.inputs[0]);
      if (it != model->initializers.end() && !model->initializers.count(n.outputs[0])) {
        OnnxTensor alias;
        alias.name = n.outputs[0];
        alias.dims = it->second.dims;
        alias.data_type = it->second.data_type;
        alias.data = it->second.data;
        alias.nbytes = it->second.nbytes;
        model->initializers.emplace(alias.name, std::move(alias));
      }
    }
  }
  // Old exporters list initializers among the graph inputs; real inputs are the ones without data.
  for (OnnxValueInfo& vi : graph_inputs) {
    if (model->initializers.count(vi.name)) continue;
    model->inputs.push_back(vi.name);
    model->input_infos.push_back(std::move(vi));
  }
  return true;
}

}  // namespace clipb200

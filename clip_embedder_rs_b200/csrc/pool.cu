// In-process multi-GPU pool behind the C ABI (include/clipb200.h, clipb200_pool_*).
//
// The reference scales out by `duplicate()`-ing an embedder (src/vision.rs:87-91, src/text.rs:104-108, src/clip.rs:69-73)
// and letting the caller spread work over the copies.  The pool is that, done once, for the GPUs of one box: one engine
// replica per device, one persistent host thread per replica, and a batch is split into contiguous row ranges — rows
// [start_r, start_r + n_r) of the caller's input go to replica r and its embeddings are copied straight into rows
// [start_r, ...) of the caller's output.  No collective: the towers are independent per row (BASELINE north_star).
// Each replica runs its own H2D / compute / D2H pipeline (Engine::RunPipelined); pinned caller buffers are used in
// place, pageable ones are staged by the replica's own thread, so the staging copies run in parallel as well.
#include <cuda_runtime.h>

#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/clipb200.h"
#include "engine.h"

using clipb200::Engine;
using clipb200::Status;

namespace {

// A worker owns one replica; jobs are closures run on the worker's thread (which has the device current).
class Worker {
 public:
  Worker() : thread_([this] { Loop(); }) {}
  ~Worker() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    thread_.join();
  }
  void Post(std::function<Status()> fn) {
    {
      std::lock_guard<std::mutex> g(mu_);
      job_ = std::move(fn);
      has_job_ = true;
      done_ = false;
    }
    cv_.notify_all();
  }
  Status Wait() {
    std::unique_lock<std::mutex> g(mu_);
    cv_.wait(g, [this] { return done_; });
    return result_;
  }

 private:
  void Loop() {
    for (;;) {
      std::function<Status()> fn;
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [this] { return has_job_ || stop_; });
        if (stop_ && !has_job_) return;
        fn = std::move(job_);
        has_job_ = false;
      }
      Status s;
      try {
        s = fn();
      } catch (const std::exception& ex) {
        s = Status::Err(CLIPB200_ERR_INVALID_ARG, std::string("exception: ") + ex.what());
      } catch (...) {
        s = Status::Err(CLIPB200_ERR_INVALID_ARG, "unknown exception");
      }
      {
        std::lock_guard<std::mutex> g(mu_);
        result_ = s;
        done_ = true;
      }
      cv_.notify_all();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_;
  std::function<Status()> job_;
  bool has_job_ = false, done_ = true, stop_ = false;
  Status result_;
  std::thread thread_;  // last member: started after everything above is initialised
};

}  // namespace

struct clipb200_pool {
  std::vector<int> devices;
  std::vector<Engine*> engines;
  std::vector<std::unique_ptr<Worker>> workers;
  ~clipb200_pool() {
    // destroy each replica on its own thread (frees device memory with the right device current), then stop the threads
    for (size_t r = 0; r < engines.size(); ++r) {
      Engine* e = engines[r];
      if (e == nullptr) continue;
      workers[r]->Post([e]() {
        delete e;
        return Status::OK();
      });
    }
    for (size_t r = 0; r < engines.size(); ++r)
      if (engines[r] != nullptr) workers[r]->Wait();
    workers.clear();
  }
};

extern int clipb200_set_last_error(int code, const std::string& msg);  // capi.cu

// contiguous ranges, the first (batch % n) replicas take one extra row — the same split as
// clip_embedder_rs_b200/sharding.py::shard_range, so a pool and N torchrun ranks see identical shards
static void shard_range(int64_t batch, int n, int r, int64_t* start, int64_t* count) {
  const int64_t base = batch / n, extra = batch % n;
  *start = r * base + (r < extra ? r : extra);
  *count = base + (r < extra ? 1 : 0);
}

template <typename Fn>
static int run_sharded(clipb200_pool* p, int64_t batch, Fn&& fn) {
  const int n = static_cast<int>(p->engines.size());
  std::vector<bool> posted(n, false);
  for (int r = 0; r < n; ++r) {
    int64_t start, count;
    shard_range(batch, n, r, &start, &count);
    if (count <= 0) continue;
    Engine* e = p->engines[r];
    p->workers[r]->Post([e, start, count, &fn]() { return fn(e, start, count); });
    posted[r] = true;
  }
  Status first;
  int first_rank = -1;
  for (int r = 0; r < n; ++r) {
    if (!posted[r]) continue;
    Status s = p->workers[r]->Wait();
    if (!s.ok() && first.ok()) {
      first = s;
      first_rank = r;
    }
  }
  if (first.ok()) return CLIPB200_OK;
  return clipb200_set_last_error(first.code, "replica " + std::to_string(first_rank) + " (cuda:" +
                                                 std::to_string(p->devices[first_rank]) + "): " + first.msg);
}

extern "C" {

int clipb200_pool_create(const char* onnx_path, const int32_t* devices, int32_t n_devices, const clipb200_opts* opts,
                         clipb200_pool** out) {
  try {
    if (onnx_path == nullptr || out == nullptr) return clipb200_set_last_error(CLIPB200_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    std::vector<int> devs;
    if (devices == nullptr || n_devices <= 0) {  // all visible devices
      int count = 0;
      if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return clipb200_set_last_error(CLIPB200_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
      }
      for (int i = 0; i < count; ++i) devs.push_back(i);
    } else {
      if (n_devices > 64) return clipb200_set_last_error(CLIPB200_ERR_INVALID_ARG, "too many devices");
      devs.assign(devices, devices + n_devices);
    }
    std::unique_ptr<clipb200_pool> p(new clipb200_pool());
    p->devices = devs;
    p->engines.assign(devs.size(), nullptr);
    for (size_t r = 0; r < devs.size(); ++r) p->workers.emplace_back(new Worker());
    // every replica loads its weights on its own thread: N file reads + uploads run concurrently
    const std::string path(onnx_path);
    clipb200_opts o = {};
    if (opts != nullptr) o = *opts;
    for (size_t r = 0; r < devs.size(); ++r) {
      Engine** slot = &p->engines[r];
      const int dev = devs[r];
      p->workers[r]->Post([slot, dev, path, o]() { return Engine::Create(path, dev, &o, slot); });
    }
    Status first;
    size_t first_rank = 0;
    for (size_t r = 0; r < devs.size(); ++r) {
      Status s = p->workers[r]->Wait();
      if (!s.ok() && first.ok()) {
        first = s;
        first_rank = r;
      }
    }
    if (!first.ok())
      return clipb200_set_last_error(first.code, "replica " + std::to_string(first_rank) + " (cuda:" +
                                                     std::to_string(devs[first_rank]) + "): " + first.msg);
    *out = p.release();
    return CLIPB200_OK;
  } catch (const std::exception& ex) {
    return clipb200_set_last_error(CLIPB200_ERR_INVALID_ARG, std::string("exception: ") + ex.what());
  } catch (...) {
    return clipb200_set_last_error(CLIPB200_ERR_INVALID_ARG, "unknown exception");
  }
}

void clipb200_pool_destroy(clipb200_pool* p) {
  try {
    delete p;
  } catch (...) {
  }
}

int clipb200_pool_size(const clipb200_pool* p) { return p == nullptr ? 0 : static_cast<int>(p->engines.size()); }
int clipb200_pool_device(const clipb200_pool* p, int replica) {
  if (p == nullptr || replica < 0 || replica >= static_cast<int>(p->devices.size())) return -1;
  return p->devices[replica];
}
int clipb200_pool_kind(const clipb200_pool* p) { return p == nullptr ? -1 : p->engines[0]->kind; }
int64_t clipb200_pool_embed_dim(const clipb200_pool* p) { return p == nullptr ? 0 : p->engines[0]->embed_dim; }
int64_t clipb200_pool_image_size(const clipb200_pool* p) { return p == nullptr ? 0 : p->engines[0]->image_size; }
int64_t clipb200_pool_context_length(const clipb200_pool* p) { return p == nullptr ? 0 : p->engines[0]->context_length; }
int clipb200_pool_num_inputs(const clipb200_pool* p) {
  return p == nullptr ? 0 : static_cast<int>(p->engines[0]->input_names.size());
}
const char* clipb200_pool_input_name(const clipb200_pool* p, int i) {
  if (p == nullptr || i < 0 || i >= static_cast<int>(p->engines[0]->input_names.size())) return nullptr;
  return p->engines[0]->input_names[i].c_str();
}
int64_t clipb200_pool_launch_count(const clipb200_pool* p) {
  int64_t n = 0;
  if (p != nullptr)
    for (const Engine* e : p->engines) n += e->launch_count;
  return n;
}

#define POOL_GUARD_BEGIN try {
#define POOL_GUARD_END                                                                                   \
  }                                                                                                      \
  catch (const std::exception& ex) {                                                                     \
    return clipb200_set_last_error(CLIPB200_ERR_INVALID_ARG, std::string("exception: ") + ex.what());    \
  }                                                                                                      \
  catch (...) {                                                                                          \
    return clipb200_set_last_error(CLIPB200_ERR_INVALID_ARG, "unknown exception");                       \
  }

int clipb200_pool_vision_embed_rgb8(clipb200_pool* p, const uint8_t* hwc, int64_t batch, int32_t width, int32_t height,
                                    const clipb200_preproc* pp, float* out) {
  POOL_GUARD_BEGIN
  if (p == nullptr) return clipb200_set_last_error(CLIPB200_ERR_INVALID_ARG, "null pool");
  if (batch <= 0) return clipb200_set_last_error(CLIPB200_ERR_INVALID_ARG, "Empty batch");
  if (hwc == nullptr || out == nullptr) return clipb200_set_last_error(CLIPB200_ERR_INVALID_ARG, "null buffer");
  const size_t px = static_cast<size_t>(width) * height * 3;
  const size_t E = static_cast<size_t>(p->engines[0]->embed_dim);
  return run_sharded(p, batch, [=](Engine* e, int64_t start, int64_t count) {
    return e->VisionEmbedRgb8(hwc + static_cast<size_t>(start) * px, count, width, height, pp, out + static_cast<size_t>(start) * E,
                              false);
  });
  POOL_GUARD_END
}

int clipb200_pool_vision_embed_rgb8_var(clipb200_pool* p, const uint8_t* const* images, const int32_t* widths,
                                        const int32_t* heights, int64_t batch, const clipb200_preproc* pp, float* out) {
  POOL_GUARD_BEGIN
  if (p == nullptr) return clipb200_set_last_error(CLIPB200_ERR_INVALID_ARG, "null pool");
  if (batch <= 0) return clipb200_set_last_error(CLIPB200_ERR_INVALID_ARG, "Empty batch");
  if (images == nullptr || widths == nullptr || heights == nullptr || out == nullptr)
    return clipb200_set_last_error(CLIPB200_ERR_INVALID_ARG, "null buffer");
  const size_t E = static_cast<size_t>(p->engines[0]->embed_dim);
  return run_sharded(p, batch, [=](Engine* e, int64_t start, int64_t count) {
    return e->VisionEmbedRgb8Var(images + start, widths + start, heights + start, count, pp, out + static_cast<size_t>(start) * E);
  });
  POOL_GUARD_END
}

int clipb200_pool_text_embed(clipb200_pool* p, const int64_t* input_ids, const int64_t* /*attention_mask_or_null*/,
                             int64_t batch, int64_t ctx, float* out) {
  POOL_GUARD_BEGIN
  if (p == nullptr) return clipb200_set_last_error(CLIPB200_ERR_INVALID_ARG, "null pool");
  if (batch <= 0) return clipb200_set_last_error(CLIPB200_ERR_INVALID_ARG, "Empty batch");
  if (input_ids == nullptr || out == nullptr) return clipb200_set_last_error(CLIPB200_ERR_INVALID_ARG, "null buffer");
  const size_t E = static_cast<size_t>(p->engines[0]->embed_dim);
  return run_sharded(p, batch, [=](Engine* e, int64_t start, int64_t count) {
    return e->TextEmbed(input_ids + static_cast<size_t>(start) * ctx, count, ctx, out + static_cast<size_t>(start) * E, false);
  });
  POOL_GUARD_END
}

}  // extern "C"

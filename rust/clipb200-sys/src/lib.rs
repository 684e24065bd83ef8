//! Rust side of the C ABI in `include/clipb200.h`.
//!
//! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Rust toolchain.  The same
//! boundary is exercised by the Python `ctypes` binding (`clip_embedder_rs_b200/_native.py`) in every GPU test.
//! `rust/open_clip_inference_b200` holds the four source files that turn the reference crate into a client of this one.
//!
//! `Engine` is what `open_clip_inference::onnx::OnnxSession` would hold instead of `RwLock<ort::Session>`
//! (upstream `src/onnx.rs:8-11`); `embed_rgb8` / `embed_pixel_values` / `embed_ids` replace the three
//! `session.run` call sites (upstream `src/vision.rs:105-113`, `src/text.rs:153-166`); `similarity` replaces the
//! ndarray tail of `src/clip.rs:102-121`.  Errors carry the engine's message the way `ClipError::Ort(String)` does.
use std::ffi::{CStr, CString};
use std::os::raw::{c_char, c_int};
use std::path::Path;

#[repr(C)]
pub struct RawEngine {
    _opaque: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct Opts {
    pub micro_batch: i32,
    pub profile: i32,
    pub reserved: [i32; 6],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct Preproc {
    pub mean: [f32; 3],
    pub std: [f32; 3],
    /// 0 bicubic (CatmullRom), 1 bilinear, 2 nearest
    pub interpolation: i32,
    /// 0 shortest (centre crop), 1 squash
    pub resize_mode: i32,
}

impl Preproc {
    /// From the strings of `open_clip_config.json`'s `preprocess_cfg`.
    pub fn new(mean: [f32; 3], std: [f32; 3], interpolation: &str, resize_mode: &str) -> Self {
        let interpolation = match interpolation {
            "bicubic" => 0,
            "bilinear" => 1,
            _ => 2,
        };
        let resize_mode = i32::from(resize_mode == "squash");
        Self { mean, std, interpolation, resize_mode }
    }
}

extern "C" {
    fn clipb200_engine_create(path: *const c_char, device: c_int, opts: *const Opts, out: *mut *mut RawEngine) -> c_int;
    fn clipb200_engine_destroy(e: *mut RawEngine);
    fn clipb200_last_error() -> *const c_char;
    fn clipb200_engine_num_inputs(e: *const RawEngine) -> c_int;
    fn clipb200_engine_input_name(e: *const RawEngine, i: c_int) -> *const c_char;
    fn clipb200_engine_embed_dim(e: *const RawEngine) -> i64;
    fn clipb200_engine_image_size(e: *const RawEngine) -> i64;
    fn clipb200_engine_context_length(e: *const RawEngine) -> i64;
    fn clipb200_vision_embed_f32(e: *mut RawEngine, nchw: *const f32, batch: i64, out: *mut f32) -> c_int;
    fn clipb200_vision_embed_rgb8(e: *mut RawEngine, hwc: *const u8, batch: i64, w: i32, h: i32, pp: *const Preproc,
                                  out: *mut f32) -> c_int;
    fn clipb200_vision_embed_rgb8_var(e: *mut RawEngine, images: *const *const u8, widths: *const i32,
                                      heights: *const i32, batch: i64, pp: *const Preproc, out: *mut f32) -> c_int;
    fn clipb200_text_embed(e: *mut RawEngine, ids: *const i64, mask: *const i64, batch: i64, ctx: i64,
                           out: *mut f32) -> c_int;
    fn clipb200_similarity(device: c_int, a: *const f32, b: *const f32, n: i64, d: i64, scale: f32, bias: f32,
                           activation: c_int, probs: *mut f32) -> c_int;
    fn clipb200_resize_rgb8(e: *mut RawEngine, image: *const u8, w: i32, h: i32, pp: *const Preproc, out: *mut u8) -> c_int;
    fn clipb200_preprocess_rgb8(e: *mut RawEngine, hwc: *const u8, batch: i64, w: i32, h: i32, pp: *const Preproc,
                                out_nchw: *mut f32) -> c_int;
    fn clipb200_onnx_inspect(path: *const c_char, json_out: *mut c_char, capacity: usize) -> c_int;
    fn clipb200_corpus_create(device: c_int, dim: i64, capacity: i64, out: *mut *mut RawCorpus) -> c_int;
    fn clipb200_corpus_destroy(c: *mut RawCorpus);
    fn clipb200_corpus_append(c: *mut RawCorpus, rows: *const f32, n: i64) -> c_int;
    fn clipb200_corpus_size(c: *const RawCorpus) -> i64;
    fn clipb200_corpus_rank(c: *mut RawCorpus, query: *const f32, scale: f32, bias: f32, activation: c_int,
                            probs: *mut f32) -> c_int;
    fn clipb200_corpus_search(c: *mut RawCorpus, queries: *const f32, n_queries: i64, k: i64, scale: f32, bias: f32,
                              activation: c_int, top_index: *mut i64, top_prob: *mut f32) -> c_int;
    fn clipb200_pool_create(path: *const c_char, devices: *const i32, n_devices: i32, opts: *const Opts,
                            out: *mut *mut RawPool) -> c_int;
    fn clipb200_pool_destroy(p: *mut RawPool);
    fn clipb200_pool_size(p: *const RawPool) -> c_int;
    fn clipb200_pool_embed_dim(p: *const RawPool) -> i64;
    fn clipb200_pool_image_size(p: *const RawPool) -> i64;
    fn clipb200_pool_context_length(p: *const RawPool) -> i64;
    fn clipb200_pool_num_inputs(p: *const RawPool) -> c_int;
    fn clipb200_pool_input_name(p: *const RawPool, i: c_int) -> *const c_char;
    fn clipb200_pool_vision_embed_rgb8(p: *mut RawPool, hwc: *const u8, batch: i64, w: i32, h: i32, pp: *const Preproc,
                                       out: *mut f32) -> c_int;
    fn clipb200_pool_vision_embed_rgb8_var(p: *mut RawPool, images: *const *const u8, widths: *const i32,
                                           heights: *const i32, batch: i64, pp: *const Preproc, out: *mut f32) -> c_int;
    fn clipb200_pool_text_embed(p: *mut RawPool, ids: *const i64, mask: *const i64, batch: i64, ctx: i64,
                                out: *mut f32) -> c_int;
}

#[repr(C)]
pub struct RawPool {
    _private: [u8; 0],
}

#[repr(C)]
pub struct RawCorpus {
    _private: [u8; 0],
}

/// Engine failure; maps onto `ClipError::Ort(String)` upstream.
#[derive(Debug, Clone)]
pub struct EngineError {
    pub code: i32,
    pub message: String,
}

impl std::fmt::Display for EngineError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "clipb200 error {}: {}", self.code, self.message)
    }
}
impl std::error::Error for EngineError {}

fn check(rc: c_int) -> Result<(), EngineError> {
    if rc == 0 {
        return Ok(());
    }
    let message = unsafe { CStr::from_ptr(clipb200_last_error()) }.to_string_lossy().into_owned();
    Err(EngineError { code: rc, message })
}

/// Every safe wrapper checks its slices against the shapes the C side will read: a short slice must be an error here,
/// never an out-of-bounds read behind the FFI boundary.
fn invalid(message: impl Into<String>) -> EngineError {
    EngineError { code: 1, message: message.into() }
}
fn expect_len(what: &str, got: usize, want: usize) -> Result<(), EngineError> {
    if got == want {
        Ok(())
    } else {
        Err(invalid(format!("{what}: slice has {got} elements, the call needs {want}")))
    }
}

/// One (model file, GPU) pair.  Not re-entrant: callers keep it behind the same `RwLock` write guard the reference
/// takes around `session.run`.
pub struct Engine {
    raw: *mut RawEngine,
}

// The handle owns device resources only; it may move between threads.  Every method that launches work takes
// `&mut self`; the `&self` methods only read fields that never change after `new`, so sharing references is sound.
unsafe impl Send for Engine {}
unsafe impl Sync for Engine {}

impl Engine {
    pub fn new(onnx_path: impl AsRef<Path>, cuda_device: i32) -> Result<Self, EngineError> {
        let path = CString::new(onnx_path.as_ref().to_string_lossy().as_bytes())
            .map_err(|e| EngineError { code: 1, message: e.to_string() })?;
        let mut raw = std::ptr::null_mut();
        check(unsafe { clipb200_engine_create(path.as_ptr(), cuda_device, &Opts::default(), &mut raw) })?;
        Ok(Self { raw })
    }

    /// `session.inputs()` names, for `has_input` / `find_input`.
    pub fn input_names(&self) -> Vec<String> {
        let n = unsafe { clipb200_engine_num_inputs(self.raw) };
        (0..n)
            .map(|i| unsafe { CStr::from_ptr(clipb200_engine_input_name(self.raw, i)) }.to_string_lossy().into_owned())
            .collect()
    }
    pub fn embed_dim(&self) -> usize {
        unsafe { clipb200_engine_embed_dim(self.raw) as usize }
    }
    pub fn image_size(&self) -> usize {
        unsafe { clipb200_engine_image_size(self.raw) as usize }
    }
    pub fn context_length(&self) -> usize {
        unsafe { clipb200_engine_context_length(self.raw) as usize }
    }

    /// `pixel_values` f32 [B,3,S,S] in, `[B, embed_dim]` L2-normalised rows out (the ORT-shaped call).
    pub fn embed_pixel_values(&mut self, nchw: &[f32], batch: usize) -> Result<Vec<f32>, EngineError> {
        let s = self.image_size();
        if batch == 0 {
            return Err(invalid("Empty batch"));
        }
        expect_len("pixel_values", nchw.len(), batch * 3 * s * s)?;
        let mut out = vec![0f32; batch * self.embed_dim()];
        check(unsafe { clipb200_vision_embed_f32(self.raw, nchw.as_ptr(), batch as i64, out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// Packed RGB8 images already at the model resolution; normalisation runs on the GPU.
    pub fn embed_rgb8(&mut self, hwc: &[u8], batch: usize, pp: &Preproc) -> Result<Vec<f32>, EngineError> {
        let s = self.image_size();
        if batch == 0 {
            return Err(invalid("Empty batch"));
        }
        expect_len("rgb8 batch", hwc.len(), batch * s * s * 3)?;
        let mut out = vec![0f32; batch * self.embed_dim()];
        check(unsafe {
            clipb200_vision_embed_rgb8(self.raw, hwc.as_ptr(), batch as i64, s as i32, s as i32, pp, out.as_mut_ptr())
        })?;
        Ok(out)
    }

    /// RGB8 images of arbitrary sizes: `(pixels, width, height)` per image; resize + crop + normalise on the GPU.
    pub fn embed_rgb8_any(&mut self, images: &[(&[u8], u32, u32)], pp: &Preproc) -> Result<Vec<f32>, EngineError> {
        let (ptrs, ws, hs) = image_table(images)?;
        let mut out = vec![0f32; images.len() * self.embed_dim()];
        check(unsafe {
            clipb200_vision_embed_rgb8_var(self.raw, ptrs.as_ptr(), ws.as_ptr(), hs.as_ptr(), images.len() as i64, pp,
                                           out.as_mut_ptr())
        })?;
        Ok(out)
    }

    /// `input_ids` i64 [B, ctx].  `mask` is accepted for call-site compatibility (src/text.rs:156-161) and must have
    /// the same shape; graphs that declare an attention_mask input are refused at load time, so it is never read.
    pub fn embed_ids(&mut self, ids: &[i64], mask: Option<&[i64]>, batch: usize) -> Result<Vec<f32>, EngineError> {
        let ctx = self.context_length();
        if batch == 0 {
            return Err(invalid("Empty batch"));
        }
        expect_len("input_ids", ids.len(), batch * ctx)?;
        if let Some(m) = mask {
            expect_len("attention_mask", m.len(), batch * ctx)?;
        }
        let mut out = vec![0f32; batch * self.embed_dim()];
        let mask_ptr = mask.map_or(std::ptr::null(), <[i64]>::as_ptr);
        check(unsafe { clipb200_text_embed(self.raw, ids.as_ptr(), mask_ptr, batch as i64, ctx as i64, out.as_mut_ptr()) })?;
        Ok(out)
    }
}

/// Pointer / width / height arrays of a list of RGB8 images, every pixel slice checked against its dimensions.
fn image_table(images: &[(&[u8], u32, u32)]) -> Result<(Vec<*const u8>, Vec<i32>, Vec<i32>), EngineError> {
    if images.is_empty() {
        return Err(invalid("Empty batch"));
    }
    let mut ptrs = Vec::with_capacity(images.len());
    let mut ws = Vec::with_capacity(images.len());
    let mut hs = Vec::with_capacity(images.len());
    for (i, (pixels, w, h)) in images.iter().enumerate() {
        if *w == 0 || *h == 0 || *w > 32768 || *h > 32768 {
            return Err(invalid(format!("image {i}: bad size {w}x{h}")));
        }
        expect_len(&format!("image {i} ({w}x{h} RGB8)"), pixels.len(), *w as usize * *h as usize * 3)?;
        ptrs.push(pixels.as_ptr());
        ws.push(*w as i32);
        hs.push(*h as i32);
    }
    Ok((ptrs, ws, hs))
}

impl Engine {
    /// `resize_with_fast_image_resize` (src/vision.rs:164-198) on the GPU: one RGB8 image in, `S x S x 3` RGB8 out.
    pub fn resize_rgb8(&mut self, pixels: &[u8], width: u32, height: u32, pp: &Preproc) -> Result<Vec<u8>, EngineError> {
        let s = self.image_size();
        image_table(&[(pixels, width, height)])?;
        let mut out = vec![0u8; s * s * 3];
        check(unsafe { clipb200_resize_rgb8(self.raw, pixels.as_ptr(), width as i32, height as i32, pp, out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// `preprocess_batch` (src/vision.rs:120-135) for images already at the model resolution: f32 [B,3,S,S], bit-exact.
    pub fn preprocess_rgb8(&mut self, hwc: &[u8], batch: usize, pp: &Preproc) -> Result<Vec<f32>, EngineError> {
        let s = self.image_size();
        if batch == 0 {
            return Err(invalid("Empty batch"));
        }
        expect_len("rgb8 batch", hwc.len(), batch * s * s * 3)?;
        let mut out = vec![0f32; batch * 3 * s * s];
        check(unsafe {
            clipb200_preprocess_rgb8(self.raw, hwc.as_ptr(), batch as i64, s as i32, s as i32, pp, out.as_mut_ptr())
        })?;
        Ok(out)
    }
}

impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { clipb200_engine_destroy(self.raw) }
    }
}

/// Parse-only description (JSON) of a model file, including how the graph recogniser bound every parameter.
pub fn inspect(onnx_path: impl AsRef<Path>) -> Result<String, EngineError> {
    let path = CString::new(onnx_path.as_ref().to_string_lossy().as_bytes())
        .map_err(|e| EngineError { code: 1, message: e.to_string() })?;
    let mut buf = vec![0u8; 1 << 22];
    check(unsafe { clipb200_onnx_inspect(path.as_ptr(), buf.as_mut_ptr().cast::<c_char>(), buf.len()) })?;
    let end = buf.iter().position(|&b| b == 0).unwrap_or(buf.len());
    Ok(String::from_utf8_lossy(&buf[..end]).into_owned())
}

/// HBM-resident embedding matrix for `rank_images` over corpora that do not fit one call (src/clip.rs:136-170).
pub struct Corpus {
    raw: *mut RawCorpus,
    dim: usize,
}

unsafe impl Send for Corpus {}

impl Corpus {
    pub fn new(cuda_device: i32, dim: usize, capacity: usize) -> Result<Self, EngineError> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { clipb200_corpus_create(cuda_device, dim as i64, capacity as i64, &mut raw) })?;
        Ok(Self { raw, dim })
    }
    pub fn len(&self) -> usize {
        unsafe { clipb200_corpus_size(self.raw) as usize }
    }
    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }
    /// Appends `rows.len() / dim` embeddings (row-major).
    pub fn append(&mut self, rows: &[f32]) -> Result<(), EngineError> {
        if rows.len() % self.dim != 0 {
            return Err(invalid(format!("corpus rows: {} values is not a multiple of dim {}", rows.len(), self.dim)));
        }
        check(unsafe { clipb200_corpus_append(self.raw, rows.as_ptr(), (rows.len() / self.dim) as i64) })
    }
    /// `rank_images` for many queries at once: one GEMM pass over the corpus and a top-k on the GPU.  Returns
    /// `(indices, probabilities)`, each `queries.len() / dim` rows of `k`, in the reference's stable descending order.
    pub fn search(&mut self, queries: &[f32], k: usize, scale: f32, bias: f32, sigmoid: bool)
                  -> Result<(Vec<i64>, Vec<f32>), EngineError> {
        if queries.is_empty() || queries.len() % self.dim != 0 {
            return Err(invalid(format!("queries: {} values is not a positive multiple of dim {}", queries.len(), self.dim)));
        }
        let nq = queries.len() / self.dim;
        if k == 0 || k > self.len() {
            return Err(invalid(format!("k = {k} outside [1, {}]", self.len())));
        }
        let mut index = vec![0i64; nq * k];
        let mut prob = vec![0f32; nq * k];
        check(unsafe {
            clipb200_corpus_search(self.raw, queries.as_ptr(), nq as i64, k as i64, scale, bias, c_int::from(sigmoid),
                                   index.as_mut_ptr(), prob.as_mut_ptr())
        })?;
        Ok((index, prob))
    }
    /// One fused similarity pass over the corpus; the stable descending sort stays with the caller (clip.rs:167).
    pub fn rank(&mut self, query: &[f32], scale: f32, bias: f32, sigmoid: bool) -> Result<Vec<f32>, EngineError> {
        expect_len("query", query.len(), self.dim)?;
        let mut probs = vec![0f32; self.len()];
        check(unsafe {
            clipb200_corpus_rank(self.raw, query.as_ptr(), scale, bias, c_int::from(sigmoid), probs.as_mut_ptr())
        })?;
        Ok(probs)
    }
}

impl Drop for Corpus {
    fn drop(&mut self) {
        unsafe { clipb200_corpus_destroy(self.raw) }
    }
}

/// `probs[i] = act(mul_add(dot(embs[i], query), scale, bias))`; softmax over all rows unless `sigmoid`.
pub fn similarity(embs: &[f32], query: &[f32], scale: f32, bias: f32, sigmoid: bool) -> Result<Vec<f32>, EngineError> {
    similarity_on(0, embs, query, scale, bias, c_int::from(sigmoid))
}

/// Raw logits `mul_add(dot, scale, bias)` (what `Clip::compare` returns, src/clip.rs:81-90).
pub fn logits(embs: &[f32], query: &[f32], scale: f32, bias: f32) -> Result<Vec<f32>, EngineError> {
    similarity_on(0, embs, query, scale, bias, 2)
}

fn similarity_on(device: i32, embs: &[f32], query: &[f32], scale: f32, bias: f32, activation: c_int)
                 -> Result<Vec<f32>, EngineError> {
    let d = query.len();
    if d == 0 || embs.is_empty() || embs.len() % d != 0 {
        return Err(invalid(format!("similarity: {} embedding values against a query of {d}", embs.len())));
    }
    let n = embs.len() / d;
    let mut probs = vec![0f32; n];
    check(unsafe {
        clipb200_similarity(device, embs.as_ptr(), query.as_ptr(), n as i64, d as i64, scale, bias, activation,
                            probs.as_mut_ptr())
    })?;
    Ok(probs)
}

/// In-process multi-GPU pool (`clipb200_pool_*`): one replica and one host thread per device, a batch is split into
/// contiguous row ranges.  What N `duplicate()`s (src/vision.rs:87-91) plus a caller-side split would do.
pub struct Pool {
    raw: *mut RawPool,
}

unsafe impl Send for Pool {}
unsafe impl Sync for Pool {}

impl Pool {
    /// `devices` empty = every visible GPU.
    pub fn new(onnx_path: impl AsRef<Path>, devices: &[i32]) -> Result<Self, EngineError> {
        let path = CString::new(onnx_path.as_ref().to_string_lossy().as_bytes()).map_err(|e| invalid(e.to_string()))?;
        let mut raw = std::ptr::null_mut();
        check(unsafe { clipb200_pool_create(path.as_ptr(), devices.as_ptr(), devices.len() as i32, &Opts::default(), &mut raw) })?;
        Ok(Self { raw })
    }
    pub fn replicas(&self) -> usize {
        unsafe { clipb200_pool_size(self.raw) as usize }
    }
    pub fn input_names(&self) -> Vec<String> {
        let n = unsafe { clipb200_pool_num_inputs(self.raw) };
        (0..n).map(|i| unsafe { CStr::from_ptr(clipb200_pool_input_name(self.raw, i)) }.to_string_lossy().into_owned()).collect()
    }
    pub fn embed_dim(&self) -> usize {
        unsafe { clipb200_pool_embed_dim(self.raw) as usize }
    }
    pub fn image_size(&self) -> usize {
        unsafe { clipb200_pool_image_size(self.raw) as usize }
    }
    pub fn context_length(&self) -> usize {
        unsafe { clipb200_pool_context_length(self.raw) as usize }
    }
    pub fn embed_rgb8(&mut self, hwc: &[u8], batch: usize, pp: &Preproc) -> Result<Vec<f32>, EngineError> {
        let s = self.image_size();
        if batch == 0 {
            return Err(invalid("Empty batch"));
        }
        expect_len("rgb8 batch", hwc.len(), batch * s * s * 3)?;
        let mut out = vec![0f32; batch * self.embed_dim()];
        check(unsafe {
            clipb200_pool_vision_embed_rgb8(self.raw, hwc.as_ptr(), batch as i64, s as i32, s as i32, pp, out.as_mut_ptr())
        })?;
        Ok(out)
    }
    pub fn embed_rgb8_any(&mut self, images: &[(&[u8], u32, u32)], pp: &Preproc) -> Result<Vec<f32>, EngineError> {
        let (ptrs, ws, hs) = image_table(images)?;
        let mut out = vec![0f32; images.len() * self.embed_dim()];
        check(unsafe {
            clipb200_pool_vision_embed_rgb8_var(self.raw, ptrs.as_ptr(), ws.as_ptr(), hs.as_ptr(), images.len() as i64, pp,
                                                out.as_mut_ptr())
        })?;
        Ok(out)
    }
    pub fn embed_ids(&mut self, ids: &[i64], batch: usize) -> Result<Vec<f32>, EngineError> {
        let ctx = self.context_length();
        if batch == 0 {
            return Err(invalid("Empty batch"));
        }
        expect_len("input_ids", ids.len(), batch * ctx)?;
        let mut out = vec![0f32; batch * self.embed_dim()];
        check(unsafe { clipb200_pool_text_embed(self.raw, ids.as_ptr(), std::ptr::null(), batch as i64, ctx as i64, out.as_mut_ptr()) })?;
        Ok(out)
    }
}

impl Drop for Pool {
    fn drop(&mut self) {
        unsafe { clipb200_pool_destroy(self.raw) }
    }
}

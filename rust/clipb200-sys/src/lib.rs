//! Rust side of the C ABI in `include/clipb200.h`.
//!
//! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Rust toolchain.  The same
//! boundary is exercised by the Python `ctypes` binding (`clip_embedder_rs_b200/_native.py`) in every GPU test.
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

/// One (model file, GPU) pair.  Not re-entrant: callers keep it behind the same `RwLock` write guard the reference
/// takes around `session.run`.
pub struct Engine {
    raw: *mut RawEngine,
}

// The handle owns device resources only; it may move between threads but is used by one thread at a time.
unsafe impl Send for Engine {}

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
        let mut out = vec![0f32; batch * self.embed_dim()];
        check(unsafe { clipb200_vision_embed_f32(self.raw, nchw.as_ptr(), batch as i64, out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// Packed RGB8 images already at the model resolution; normalisation runs on the GPU.
    pub fn embed_rgb8(&mut self, hwc: &[u8], batch: usize, pp: &Preproc) -> Result<Vec<f32>, EngineError> {
        let s = self.image_size() as i32;
        let mut out = vec![0f32; batch * self.embed_dim()];
        check(unsafe { clipb200_vision_embed_rgb8(self.raw, hwc.as_ptr(), batch as i64, s, s, pp, out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// RGB8 images of arbitrary sizes: `(pixels, width, height)` per image; resize + crop + normalise on the GPU.
    pub fn embed_rgb8_any(&mut self, images: &[(&[u8], u32, u32)], pp: &Preproc) -> Result<Vec<f32>, EngineError> {
        let ptrs: Vec<*const u8> = images.iter().map(|(p, _, _)| p.as_ptr()).collect();
        let ws: Vec<i32> = images.iter().map(|(_, w, _)| *w as i32).collect();
        let hs: Vec<i32> = images.iter().map(|(_, _, h)| *h as i32).collect();
        let mut out = vec![0f32; images.len() * self.embed_dim()];
        check(unsafe {
            clipb200_vision_embed_rgb8_var(self.raw, ptrs.as_ptr(), ws.as_ptr(), hs.as_ptr(), images.len() as i64, pp,
                                           out.as_mut_ptr())
        })?;
        Ok(out)
    }

    /// `input_ids` i64 [B, ctx] (and the optional mask the reference passes when the graph declares it).
    pub fn embed_ids(&mut self, ids: &[i64], mask: Option<&[i64]>, batch: usize) -> Result<Vec<f32>, EngineError> {
        let ctx = self.context_length();
        let mut out = vec![0f32; batch * self.embed_dim()];
        let mask_ptr = mask.map_or(std::ptr::null(), <[i64]>::as_ptr);
        check(unsafe { clipb200_text_embed(self.raw, ids.as_ptr(), mask_ptr, batch as i64, ctx as i64, out.as_mut_ptr()) })?;
        Ok(out)
    }
}

impl Engine {
    /// `resize_with_fast_image_resize` (src/vision.rs:164-198) on the GPU: one RGB8 image in, `S x S x 3` RGB8 out.
    pub fn resize_rgb8(&mut self, pixels: &[u8], width: u32, height: u32, pp: &Preproc) -> Result<Vec<u8>, EngineError> {
        let s = self.image_size();
        let mut out = vec![0u8; s * s * 3];
        check(unsafe { clipb200_resize_rgb8(self.raw, pixels.as_ptr(), width as i32, height as i32, pp, out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// `preprocess_batch` (src/vision.rs:120-135) for images already at the model resolution: f32 [B,3,S,S], bit-exact.
    pub fn preprocess_rgb8(&mut self, hwc: &[u8], batch: usize, pp: &Preproc) -> Result<Vec<f32>, EngineError> {
        let s = self.image_size();
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
        check(unsafe { clipb200_corpus_append(self.raw, rows.as_ptr(), (rows.len() / self.dim) as i64) })
    }
    /// One fused similarity pass over the corpus; the stable descending sort stays with the caller (clip.rs:167).
    pub fn rank(&mut self, query: &[f32], scale: f32, bias: f32, sigmoid: bool) -> Result<Vec<f32>, EngineError> {
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
    let d = query.len();
    let n = embs.len() / d;
    let mut probs = vec![0f32; n];
    check(unsafe {
        clipb200_similarity(0, embs.as_ptr(), query.as_ptr(), n as i64, d as i64, scale, bias, c_int::from(sigmoid),
                            probs.as_mut_ptr())
    })?;
    Ok(probs)
}

//! `OnnxSession` with the `ort::Session` swapped for a handle to the B200 engine (`libclipb200.so`, C ABI in
//! `include/clipb200.h`, Rust binding `clipb200-sys`).
//!
//! Replaces upstream `src/onnx.rs`.  What stays: the type name, the two public fields, `new`, `has_input`,
//! `find_input`, the `RwLock` the embedders take a write guard on around every run.  What changes: the lock guards an
//! engine (or an in-process multi-GPU pool of engines) instead of an ONNX Runtime session, and the execution-provider
//! list names CUDA devices instead of `ort` providers.  NOT COMPILED in this repository (no Rust toolchain in the build
//! image); see README.md next to Cargo.toml for the `cargo test` recipe.
use crate::ClipError;
use clipb200_sys::{Engine, EngineError, Pool, Preproc};
use std::path::Path;
use std::sync::RwLock;

/// Stands where `ort::ep::ExecutionProviderDispatch` stood in the builders' `with_execution_providers(&[...])`.
/// An empty list means "GPU 0".  One `Device` selects that GPU; several devices (or `AllDevices`) make the session a
/// pool: one replica per GPU, every batch split into contiguous row ranges (upstream's `duplicate()` done once).
#[derive(Debug, Clone, PartialEq, Eq)]
pub enum ExecutionProviderDispatch {
    Device(i32),
    AllDevices,
}

/// What the lock guards: the thing `session.run` is called on.
#[derive(Debug)]
pub enum Backend {
    Single(EngineBox),
    Pool(PoolBox),
}

/// Newtypes so that the public field keeps a `Debug` impl like `ort::Session` has.
pub struct EngineBox(pub Engine);
pub struct PoolBox(pub Pool);
impl std::fmt::Debug for EngineBox {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "clipb200 engine (embed_dim {})", self.0.embed_dim())
    }
}
impl std::fmt::Debug for PoolBox {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "clipb200 pool ({} replicas, embed_dim {})", self.0.replicas(), self.0.embed_dim())
    }
}

#[derive(Debug)]
pub struct OnnxSession {
    pub session: RwLock<Backend>,
    pub execution_providers: Vec<ExecutionProviderDispatch>,
}

/// Same variant upstream maps `ort::Error` to; "Empty batch" keeps upstream's `Inference` kind.
impl From<EngineError> for ClipError {
    fn from(err: EngineError) -> Self {
        if err.code == 1 && err.message == "Empty batch" {
            Self::Inference(err.message)
        } else {
            Self::Ort(err.message)
        }
    }
}

impl OnnxSession {
    pub fn new(
        path: impl AsRef<Path>,
        execution_providers: &[ExecutionProviderDispatch],
    ) -> Result<Self, ClipError> {
        let mut devices: Vec<i32> = Vec::new();
        let mut all = false;
        for ep in execution_providers {
            match ep {
                ExecutionProviderDispatch::Device(d) => devices.push(*d),
                ExecutionProviderDispatch::AllDevices => all = true,
            }
        }
        let backend = if all {
            Backend::Pool(PoolBox(Pool::new(path.as_ref(), &[])?))
        } else if devices.len() > 1 {
            Backend::Pool(PoolBox(Pool::new(path.as_ref(), &devices)?))
        } else {
            Backend::Single(EngineBox(Engine::new(path.as_ref(), devices.first().copied().unwrap_or(0))?))
        };
        Ok(Self {
            session: RwLock::new(backend),
            execution_providers: execution_providers.to_vec(),
        })
    }

    fn input_names(&self) -> Result<Vec<String>, ClipError> {
        let guard = self.session.read()?;
        Ok(match &*guard {
            Backend::Single(e) => e.0.input_names(),
            Backend::Pool(p) => p.0.input_names(),
        })
    }

    /// Does the model declare an input of this name?
    pub fn has_input(&self, name: &str) -> Result<bool, ClipError> {
        Ok(self.input_names()?.iter().any(|n| n == name))
    }

    /// First of `possibilities` that the model declares as an input.
    pub fn find_input(&self, possibilities: &[&str]) -> Result<Option<String>, ClipError> {
        let names = self.input_names()?;
        Ok(possibilities
            .iter()
            .find(|p| names.iter().any(|n| n == *p))
            .map(|p| (*p).to_string()))
    }

    pub fn embed_dim(&self) -> Result<usize, ClipError> {
        let guard = self.session.read()?;
        Ok(match &*guard {
            Backend::Single(e) => e.0.embed_dim(),
            Backend::Pool(p) => p.0.embed_dim(),
        })
    }

    /// `session.run(inputs![pixel_values])` for RGB8 images of any size: resize, normalise and tower on the GPU.
    pub fn run_rgb8_any(&self, images: &[(&[u8], u32, u32)], pp: &Preproc) -> Result<Vec<f32>, ClipError> {
        let mut guard = self.session.write()?;
        Ok(match &mut *guard {
            Backend::Single(e) => e.0.embed_rgb8_any(images, pp)?,
            Backend::Pool(p) => p.0.embed_rgb8_any(images, pp)?,
        })
    }

    /// Same for a packed batch that is already at the model resolution.
    pub fn run_rgb8(&self, hwc: &[u8], batch: usize, pp: &Preproc) -> Result<Vec<f32>, ClipError> {
        let mut guard = self.session.write()?;
        Ok(match &mut *guard {
            Backend::Single(e) => e.0.embed_rgb8(hwc, batch, pp)?,
            Backend::Pool(p) => p.0.embed_rgb8(hwc, batch, pp)?,
        })
    }

    /// `session.run(inputs![input_ids])`.
    pub fn run_ids(&self, ids: &[i64], batch: usize) -> Result<Vec<f32>, ClipError> {
        let mut guard = self.session.write()?;
        Ok(match &mut *guard {
            Backend::Single(e) => e.0.embed_ids(ids, None, batch)?,
            Backend::Pool(p) => p.0.embed_ids(ids, batch)?,
        })
    }

    /// Host-visible preprocessing (`preprocess_batch`): resize + normalise on the GPU, f32 NCHW back.  Single-engine
    /// sessions only (a pool has no use for it: its embed calls preprocess on every replica).
    pub fn run_preprocess(&self, images: &[(&[u8], u32, u32)], pp: &Preproc) -> Result<Vec<f32>, ClipError> {
        let mut guard = self.session.write()?;
        let Backend::Single(e) = &mut *guard else {
            return Err(ClipError::Config("preprocess_batch needs a single-device session".into()));
        };
        let size = e.0.image_size();
        let mut packed = Vec::with_capacity(images.len() * size * size * 3);
        for (pixels, w, h) in images {
            if *w as usize == size && *h as usize == size {
                packed.extend_from_slice(pixels);
            } else {
                packed.extend_from_slice(&e.0.resize_rgb8(pixels, *w, *h, pp)?);
            }
        }
        Ok(e.0.preprocess_rgb8(&packed, images.len(), pp)?)
    }
}

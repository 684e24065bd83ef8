//! `VisionEmbedder` on the B200 engine.  Replaces upstream `src/vision.rs`.
//!
//! Public surface as upstream: the struct and its five public fields, the `from_hf` / `from_local_id` /
//! `from_local_dir` builders (bon, `finish_fn = build`, optional `with_execution_providers`), `duplicate`,
//! `embed_image(s)`, `preprocess(_batch)`.  What is gone is every line of host-side arithmetic: the
//! fast_image_resize call, the crop box, the per-pixel normalisation loop and the `session.run` plumbing all live in
//! `libclipb200.so` now (resize bit-compatible with fast_image_resize's U8x3 convolution, normalisation bit-identical
//! to `(v / 255 - mean) / std`).  A `DynamicImage` is handed over as the RGB8 buffer `to_rgb8()` yields.
use crate::config::{ModelConfig, OpenClipConfig};
use crate::error::ClipError;
use crate::model_manager;
use crate::onnx::{ExecutionProviderDispatch, OnnxSession};
use bon::bon;
use clipb200_sys::Preproc;
use image::{DynamicImage, RgbImage};
use ndarray::{Array1, Array2, Array4};
use rayon::prelude::*;
use std::path::{Path, PathBuf};

#[derive(Debug)]
pub struct VisionEmbedder {
    pub session: OnnxSession,
    pub config: OpenClipConfig,
    pub model_config: ModelConfig,
    pub input_name: String,
    pub model_dir: PathBuf,
}

#[bon]
impl VisionEmbedder {
    /// From a `HuggingFace` repository that holds an exported model directory.
    #[builder(finish_fn = build)]
    #[cfg(feature = "hf-hub")]
    pub async fn from_hf(
        #[builder(start_fn)] model_id: &str,
        with_execution_providers: Option<&[ExecutionProviderDispatch]>,
    ) -> Result<Self, ClipError> {
        let dir = model_manager::get_hf_model(model_id).await?;
        Self::open(&dir, with_execution_providers.unwrap_or_default())
    }

    /// From `<base_folder>/<model_id>` (default base folder: where `pull_onnx.py` exports to).
    #[builder(finish_fn = build)]
    pub fn from_local_id(
        #[builder(start_fn)] model_id: &str,
        base_folder: Option<&Path>,
        with_execution_providers: Option<&[ExecutionProviderDispatch]>,
    ) -> Result<Self, ClipError> {
        let base = match base_folder {
            Some(p) => p.to_path_buf(),
            None => model_manager::get_default_base_folder(),
        };
        Self::open(&base.join(model_id), with_execution_providers.unwrap_or_default())
    }

    /// From a model directory.
    #[builder(finish_fn = build)]
    pub fn from_local_dir(
        #[builder(start_fn)] model_dir: &Path,
        with_execution_providers: Option<&[ExecutionProviderDispatch]>,
    ) -> Result<Self, ClipError> {
        Self::open(model_dir, with_execution_providers.unwrap_or_default())
    }
}

impl VisionEmbedder {
    fn open(model_dir: &Path, eps: &[ExecutionProviderDispatch]) -> Result<Self, ClipError> {
        model_manager::verify_model_dir(model_dir)?;
        let session = OnnxSession::new(model_dir.join("visual.onnx"), eps)?;
        let config = OpenClipConfig::from_file(model_dir.join("open_clip_config.json"))?;
        let model_config = ModelConfig::from_file(model_dir.join("model_config.json"))?;
        let Some(input_name) = session.find_input(&["pixel_values", "input"])? else {
            return Err(ClipError::Config("Could not find vision input node".to_string()));
        };
        Ok(Self { session, config, model_config, input_name, model_dir: model_dir.to_path_buf() })
    }

    /// A second, independent instance of the same model (its own engine, streams and workspace).
    pub fn duplicate(&self) -> Result<Self, ClipError> {
        Self::open(&self.model_dir, &self.session.execution_providers)
    }

    fn preproc(&self) -> Preproc {
        let pc = &self.config.preprocess_cfg;
        Preproc::new(pc.mean, pc.std, &pc.interpolation, &pc.resize_mode)
    }

    /// `image.to_rgb8()` for the whole batch, in parallel (the only per-image host work left).
    fn to_rgb(images: &[DynamicImage]) -> Result<Vec<RgbImage>, ClipError> {
        if images.is_empty() {
            return Err(ClipError::Inference("Empty batch".to_string()));
        }
        Ok(images.par_iter().map(DynamicImage::to_rgb8).collect())
    }

    pub fn embed_image(&self, image: &DynamicImage) -> Result<Array1<f32>, ClipError> {
        let rows = self.embed_images(std::slice::from_ref(image))?;
        let n = rows.len();
        Ok(rows.into_shape_with_order(n)?)
    }

    /// `[images.len(), embed_dim]`, rows L2-normalised.
    pub fn embed_images(&self, images: &[DynamicImage]) -> Result<Array2<f32>, ClipError> {
        let rgb = Self::to_rgb(images)?;
        let size = self.config.model_cfg.vision_cfg.image_size;
        let pp = self.preproc();
        let flat = if rgb.iter().all(|im| im.width() == size && im.height() == size) {
            // already at the model resolution: one packed buffer, no resize pass
            let px = (size as usize).pow(2) * 3;
            let mut packed = Vec::with_capacity(rgb.len() * px);
            for im in &rgb {
                packed.extend_from_slice(im.as_raw());
            }
            self.session.run_rgb8(&packed, rgb.len(), &pp)?
        } else {
            let table: Vec<(&[u8], u32, u32)> =
                rgb.iter().map(|im| (im.as_raw().as_slice(), im.width(), im.height())).collect();
            self.session.run_rgb8_any(&table, &pp)?
        };
        let dim = flat.len() / rgb.len();
        Ok(Array2::from_shape_vec((rgb.len(), dim), flat)?)
    }

    /// Normalised `pixel_values` `[B, 3, S, S]` as upstream returns them (computed on the GPU).
    pub fn preprocess_batch(&self, images: &[DynamicImage]) -> Result<Array4<f32>, ClipError> {
        let rgb = Self::to_rgb(images)?;
        let size = self.config.model_cfg.vision_cfg.image_size as usize;
        let table: Vec<(&[u8], u32, u32)> =
            rgb.iter().map(|im| (im.as_raw().as_slice(), im.width(), im.height())).collect();
        let flat = self.session.run_preprocess(&table, &self.preproc())?;
        Ok(Array4::from_shape_vec((rgb.len(), 3, size, size), flat)?)
    }

    pub fn preprocess(&self, image: &DynamicImage) -> Result<Array4<f32>, ClipError> {
        self.preprocess_batch(std::slice::from_ref(image))
    }
}

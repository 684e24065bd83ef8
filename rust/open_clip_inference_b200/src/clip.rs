//! `Clip`: both towers plus the similarity tail.  Replaces upstream `src/clip.rs`.
//!
//! Same struct, builders and methods.  `classify` / `rank_images` / `compare` send the embeddings through the engine's
//! similarity kernel (`dot -> mul_add(logit_scale, logit_bias) -> softmax | sigmoid`, one warp per row); the stable
//! descending sort by probability stays here, with upstream's comparator.
use crate::config::ModelConfig;
use crate::error::ClipError;
use crate::model_manager;
use crate::onnx::ExecutionProviderDispatch;
use crate::text::TextEmbedder;
use crate::vision::VisionEmbedder;
use bon::bon;
use image::DynamicImage;
use std::cmp::Ordering;
use std::path::{Path, PathBuf};

#[derive(Debug)]
pub struct Clip {
    pub vision: VisionEmbedder,
    pub text: TextEmbedder,
    pub model_dir: PathBuf,
}

#[bon]
impl Clip {
    #[cfg(feature = "hf-hub")]
    #[builder(finish_fn = build)]
    pub async fn from_hf(
        #[builder(start_fn)] model_id: &str,
        with_execution_providers: Option<&[ExecutionProviderDispatch]>,
    ) -> Result<Self, ClipError> {
        let dir = model_manager::get_hf_model(model_id).await?;
        Self::open(&dir, with_execution_providers)
    }

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
        Self::open(&base.join(model_id), with_execution_providers)
    }

    #[builder(finish_fn = build)]
    pub fn from_local_dir(
        #[builder(start_fn)] model_dir: &Path,
        with_execution_providers: Option<&[ExecutionProviderDispatch]>,
    ) -> Result<Self, ClipError> {
        Self::open(model_dir, with_execution_providers)
    }
}

/// Probabilities in the order given -> `(item, probability)` pairs, best first; equal (or incomparable) probabilities
/// keep their input order, as `sort_by(|a, b| b.1.partial_cmp(&a.1).unwrap_or(Equal))` does upstream.
fn ranked<I>(items: impl IntoIterator<Item = I>, probs: Vec<f32>) -> Vec<(I, f32)> {
    let mut pairs: Vec<(I, f32)> = items.into_iter().zip(probs).collect();
    pairs.sort_by(|a, b| b.1.partial_cmp(&a.1).unwrap_or(Ordering::Equal));
    pairs
}

impl Clip {
    fn open(model_dir: &Path, eps: Option<&[ExecutionProviderDispatch]>) -> Result<Self, ClipError> {
        model_manager::verify_model_dir(model_dir)?;
        let vision = VisionEmbedder::from_local_dir(model_dir).maybe_with_execution_providers(eps).build()?;
        let text = TextEmbedder::from_local_dir(model_dir).maybe_with_execution_providers(eps).build()?;
        Ok(Self { vision, text, model_dir: model_dir.to_path_buf() })
    }

    pub fn duplicate(&self) -> Result<Self, ClipError> {
        Self::open(&self.model_dir, Some(&self.vision.session.execution_providers))
    }

    pub fn get_model_config(&self) -> ModelConfig {
        self.text.model_config.clone()
    }

    fn scale_bias(&self) -> (f32, f32) {
        let mc = &self.text.model_config;
        (mc.logit_scale.unwrap_or(1.0), mc.logit_bias.unwrap_or(0.0))
    }

    /// `rows` (row-major `[n, dim]`) against one query: probabilities over the n rows.
    fn probabilities(&self, rows: &[f32], query: &[f32]) -> Result<Vec<f32>, ClipError> {
        let (scale, bias) = self.scale_bias();
        let sigmoid = self.text.model_config.activation_function.as_deref().unwrap_or("softmax") == "sigmoid";
        Ok(clipb200_sys::similarity(rows, query, scale, bias, sigmoid)?)
    }

    /// Raw logit of one image against one text: `dot.mul_add(logit_scale, logit_bias)`.
    pub fn compare(&self, image: &DynamicImage, text: &str) -> Result<f32, ClipError> {
        let v = self.vision.embed_image(image)?;
        let t = self.text.embed_text(text)?;
        let (scale, bias) = self.scale_bias();
        let logit = clipb200_sys::logits(v.as_slice().expect("contiguous"), t.as_slice().expect("contiguous"), scale, bias)?;
        Ok(logit[0])
    }

    /// `(label, probability)` pairs, most probable first.
    pub fn classify<T: AsRef<str>>(&self, image: &DynamicImage, labels: &[T]) -> Result<Vec<(String, f32)>, ClipError> {
        let v = self.vision.embed_image(image)?;
        let t = self.text.embed_texts(labels)?;
        let probs = self.probabilities(t.as_slice().expect("contiguous"), v.as_slice().expect("contiguous"))?;
        Ok(ranked(labels.iter().map(|l| l.as_ref().to_string()), probs))
    }

    /// `(image index, probability)` pairs, most probable first.
    pub fn rank_images(&self, images: &[DynamicImage], text: &str) -> Result<Vec<(usize, f32)>, ClipError> {
        let v = self.vision.embed_images(images)?;
        let t = self.text.embed_text(text)?;
        let probs = self.probabilities(v.as_slice().expect("contiguous"), t.as_slice().expect("contiguous"))?;
        Ok(ranked(0..images.len(), probs))
    }

    /// Max-subtracted softmax (host arithmetic: these two are pure functions of their arguments upstream too).
    #[must_use]
    pub fn softmax(logits: &[f32]) -> Vec<f32> {
        let top = logits.iter().copied().fold(f32::NEG_INFINITY, f32::max);
        let exps: Vec<f32> = logits.iter().map(|&x| (x - top).exp()).collect();
        let total: f32 = exps.iter().sum();
        exps.into_iter().map(|e| e / total).collect()
    }

    #[must_use]
    pub fn sigmoid(logit: f32) -> f32 {
        1.0 / (1.0 + (-logit).exp())
    }
}

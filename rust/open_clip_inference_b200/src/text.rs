//! `TextEmbedder` on the B200 engine.  Replaces upstream `src/text.rs`.
//!
//! The tokenizer is the same `tokenizers` crate with the same settings (fixed-length right padding to the context
//! length with `pad_id`, truncation at the context length, optional lower-casing), so `tokenize` returns the ids and
//! the mask upstream returns; `embed_texts` hands the ids to the engine instead of `session.run`.
use crate::config::{ModelConfig, OpenClipConfig};
use crate::error::ClipError;
use crate::model_manager;
use crate::onnx::{ExecutionProviderDispatch, OnnxSession};
use bon::bon;
use ndarray::{Array1, Array2};
use std::path::{Path, PathBuf};
use tokenizers::{PaddingParams, PaddingStrategy, Tokenizer, TruncationParams};

#[derive(Debug)]
pub struct TextEmbedder {
    pub session: OnnxSession,
    pub config: OpenClipConfig,
    pub model_config: ModelConfig,
    pub model_dir: PathBuf,
    tokenizer: Tokenizer,
    id_name: String,
    mask_name: Option<String>,
}

#[bon]
impl TextEmbedder {
    #[builder(finish_fn = build)]
    #[cfg(feature = "hf-hub")]
    pub async fn from_hf(
        #[builder(start_fn)] model_id: &str,
        with_execution_providers: Option<&[ExecutionProviderDispatch]>,
    ) -> Result<Self, ClipError> {
        let dir = model_manager::get_hf_model(model_id).await?;
        Self::open(&dir, with_execution_providers.unwrap_or_default())
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
        Self::open(&base.join(model_id), with_execution_providers.unwrap_or_default())
    }

    #[builder(finish_fn = build)]
    pub fn from_local_dir(
        #[builder(start_fn)] model_dir: &Path,
        with_execution_providers: Option<&[ExecutionProviderDispatch]>,
    ) -> Result<Self, ClipError> {
        Self::open(model_dir, with_execution_providers.unwrap_or_default())
    }
}

/// Pads every encoding to exactly `ctx_len` ids with `pad_id` and cuts longer ones there.
fn fixed_length(tokenizer: &mut Tokenizer, ctx_len: usize, pad_id: u32) -> Result<(), ClipError> {
    let padding = PaddingParams { strategy: PaddingStrategy::Fixed(ctx_len), pad_id, ..Default::default() };
    let truncation = TruncationParams { max_length: ctx_len, ..Default::default() };
    tokenizer.with_padding(Some(padding)).with_truncation(Some(truncation))?;
    Ok(())
}

impl TextEmbedder {
    fn open(model_dir: &Path, eps: &[ExecutionProviderDispatch]) -> Result<Self, ClipError> {
        model_manager::verify_model_dir(model_dir)?;
        let model_config = ModelConfig::from_file(model_dir.join("model_config.json"))?;
        let session = OnnxSession::new(model_dir.join("text.onnx"), eps)?;
        let config = OpenClipConfig::from_file(model_dir.join("open_clip_config.json"))?;
        let mut tokenizer = Tokenizer::from_file(model_dir.join("tokenizer.json"))?;
        let pad_id = match model_config.pad_id {
            Some(id) => id,
            None => tokenizer
                .get_vocab(true)
                .get("<pad>")
                .copied()
                .ok_or_else(|| ClipError::Config("No pad token found in tokenizer".into()))?,
        };
        fixed_length(&mut tokenizer, config.model_cfg.text_cfg.context_length, pad_id)?;
        let Some(id_name) = session.find_input(&["input_ids"])? else {
            return Err(ClipError::Config("Could not find text input node".into()));
        };
        // The engine refuses graphs that declare an attention_mask input (it cannot know what their nodes do with it),
        // so this is None for every model that loads; the field is kept because upstream has it.
        let mask_name = session.find_input(&["attention_mask"])?;
        Ok(Self { session, config, model_config, model_dir: model_dir.to_path_buf(), tokenizer, id_name, mask_name })
    }

    pub fn duplicate(&self) -> Result<Self, ClipError> {
        Self::open(&self.model_dir, &self.session.execution_providers)
    }

    /// `(input_ids, attention_mask)`, both `[texts.len(), context_length]` i64.
    pub fn tokenize<T: AsRef<str>>(&self, texts: &[T]) -> Result<(Array2<i64>, Array2<i64>), ClipError> {
        let lower = self.model_config.tokenizer_needs_lowercase;
        let inputs: Vec<String> =
            texts.iter().map(|t| if lower { t.as_ref().to_lowercase() } else { t.as_ref().to_string() }).collect();
        let encodings = self.tokenizer.encode_batch(inputs, true)?;
        let ctx = self.config.model_cfg.text_cfg.context_length;
        let mut ids = Vec::with_capacity(encodings.len() * ctx);
        let mut mask = Vec::with_capacity(encodings.len() * ctx);
        for e in &encodings {
            ids.extend(e.get_ids().iter().map(|&x| i64::from(x)));
            mask.extend(e.get_attention_mask().iter().map(|&x| i64::from(x)));
        }
        let shape = (encodings.len(), ctx);
        Ok((Array2::from_shape_vec(shape, ids)?, Array2::from_shape_vec(shape, mask)?))
    }

    pub fn embed_text(&self, text: &str) -> Result<Array1<f32>, ClipError> {
        let rows = self.embed_texts(&[text])?;
        let n = rows.len();
        Ok(rows.into_shape_with_order(n)?)
    }

    /// `[texts.len(), embed_dim]`, rows L2-normalised.
    pub fn embed_texts<T: AsRef<str>>(&self, texts: &[T]) -> Result<Array2<f32>, ClipError> {
        if texts.is_empty() {
            return Err(ClipError::Inference("Empty batch".to_string()));
        }
        let (ids, _mask) = self.tokenize(texts)?;
        debug_assert!(self.mask_name.is_none() && !self.id_name.is_empty());
        let n = ids.nrows();
        let ids = ids.as_standard_layout();
        let flat = self.session.run_ids(ids.as_slice().expect("standard layout"), n)?;
        let dim = flat.len() / n;
        Ok(Array2::from_shape_vec((n, dim), flat)?)
    }
}

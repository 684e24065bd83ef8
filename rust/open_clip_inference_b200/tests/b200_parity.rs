//! What upstream's `tests/integration_test.rs` checks (classify an image against three labels, probabilities ordered
//! and normalised), without the network: the model directory is one written by `tools/export_synthetic.py`
//! (`CLIPB200_TEST_MODEL_DIR`, e.g. the `tiny_clip` or `vit_b32` config) and the expected values are the JSON the CPU
//! oracle prints for the same directory (`python tools/rust_expectations.py <dir> > expect.json`,
//! `CLIPB200_TEST_EXPECT`).  Needs a B200 and `libclipb200.so`; NOT RUN in this repository (no Rust toolchain here).
use open_clip_inference::onnx::ExecutionProviderDispatch;
use open_clip_inference::{Clip, ClipError, VisionEmbedder};
use std::path::PathBuf;

fn model_dir() -> Option<PathBuf> {
    std::env::var_os("CLIPB200_TEST_MODEL_DIR").map(PathBuf::from)
}

/// The counter-free deterministic test image of `tools/rust_expectations.py`: pixel (x, y, c) = (7 x + 13 y + 101 c) % 256.
fn test_image(w: u32, h: u32) -> image::DynamicImage {
    let mut buf = image::RgbImage::new(w, h);
    for (x, y, px) in buf.enumerate_pixels_mut() {
        for c in 0..3u32 {
            px.0[c as usize] = ((7 * x + 13 * y + 101 * c) % 256) as u8;
        }
    }
    image::DynamicImage::ImageRgb8(buf)
}

#[test]
fn classify_matches_the_cpu_oracle() -> Result<(), ClipError> {
    let Some(dir) = model_dir() else {
        eprintln!("CLIPB200_TEST_MODEL_DIR not set: skipping");
        return Ok(());
    };
    let clip = Clip::from_local_dir(&dir).build()?;
    let labels = ["a photo of a cat", "a photo of a dog", "a photo of a beignet"];
    let img = test_image(640, 480);
    let got = clip.classify(&img, &labels)?;
    assert_eq!(got.len(), 3);
    assert!(got.windows(2).all(|w| w[0].1 >= w[1].1), "sorted by probability");
    if clip.get_model_config().activation_function.as_deref() != Some("sigmoid") {
        assert!((got.iter().map(|(_, p)| p).sum::<f32>() - 1.0).abs() < 1e-4);
    }
    if let Some(path) = std::env::var_os("CLIPB200_TEST_EXPECT") {
        let want: serde_json::Value = serde_json::from_str(&std::fs::read_to_string(path)?)?;
        let order: Vec<String> = want["classify"].as_array().unwrap().iter().map(|e| e[0].as_str().unwrap().to_string()).collect();
        assert_eq!(got.iter().map(|(l, _)| l.clone()).collect::<Vec<_>>(), order, "same label order as the oracle");
        for (g, w) in got.iter().zip(want["classify"].as_array().unwrap()) {
            assert!((g.1 - w[1].as_f64().unwrap() as f32).abs() < 3e-2);
        }
        let emb = clip.vision.embed_image(&img)?;
        let want_emb: Vec<f32> = want["image_embedding"].as_array().unwrap().iter().map(|v| v.as_f64().unwrap() as f32).collect();
        let cos: f32 = emb.iter().zip(&want_emb).map(|(a, b)| a * b).sum();
        assert!(cos >= 0.999, "cosine vs the oracle {cos}");
    }
    Ok(())
}

#[test]
fn embeddings_are_unit_rows_and_batches_are_consistent() -> Result<(), ClipError> {
    let Some(dir) = model_dir() else { return Ok(()) };
    let vision = VisionEmbedder::from_local_dir(&dir).build()?;
    let imgs = vec![test_image(640, 480), test_image(300, 500), test_image(1944, 2592)];
    let all = vision.embed_images(&imgs)?;
    assert_eq!(all.nrows(), 3);
    for (i, row) in all.rows().into_iter().enumerate() {
        assert!((row.dot(&row).sqrt() - 1.0).abs() < 1e-3);
        let single = vision.embed_image(&imgs[i])?;
        assert!(row.iter().zip(single.iter()).all(|(a, b)| (a - b).abs() < 2e-3), "batch position must not matter");
    }
    assert!(matches!(vision.embed_images(&[]), Err(ClipError::Inference(_))));
    // the same model as an in-process pool over every GPU of the box gives the same rows
    let pool = VisionEmbedder::from_local_dir(&dir).with_execution_providers(&[ExecutionProviderDispatch::AllDevices]).build()?;
    let pooled = pool.embed_images(&imgs)?;
    assert_eq!(pooled, all);
    let twin = vision.duplicate()?;
    assert_eq!(twin.embed_images(&imgs)?, all);
    Ok(())
}

//! imagekit-cuda: `resize_image` with the reference's exact signature
//! (imagekit src/transform.rs:62-66), executed on B200 GPUs through libimagekit_cuda.so.
//!
//! Drop-in: in imagekit's src/transform.rs replace the body of `resize_image` with
//! `imagekit_cuda::resize_image(img, w, h).map_err(ImageKitError::TransformError)`; the handlers
//! (src/lib.rs:180, :286), DiskCache and the signing path are untouched.
pub mod ffi;

use image::{DynamicImage, GenericImageView, ImageBuffer};
use std::ffi::CStr;
use std::sync::OnceLock;

struct Ctx(*mut ffi::ikc_ctx);
unsafe impl Send for Ctx {} // the C ABI is thread-safe and re-entrant
unsafe impl Sync for Ctx {}

static CTX: OnceLock<Result<Ctx, String>> = OnceLock::new();

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::ikc_last_error()).to_string_lossy().into_owned() }
}

fn ctx() -> Result<*mut ffi::ikc_ctx, String> {
    CTX.get_or_init(|| {
        let mut p = std::ptr::null_mut();
        // all visible devices; single-image calls are spread round-robin, batches are sharded
        let rc = unsafe { ffi::ikc_create(std::ptr::null(), 0, &mut p) };
        if rc != ffi::IKC_OK {
            return Err(last_error());
        }
        // Arithmetic mode.  Default FAST: fused single-launch kernels, every u8 sample within +-1 of the CPU crate's
        // result (the drop-in's stated tolerance).  IMAGEKIT_CUDA_MODE=exact selects the two-launch kernels that
        // reproduce image 0.25.8's f32 arithmetic operation by operation (bit-identical output, several times slower).
        if std::env::var("IMAGEKIT_CUDA_MODE").map(|v| v.eq_ignore_ascii_case("exact")).unwrap_or(false) {
            unsafe { ffi::ikc_set_mode(p, ffi::IKC_MODE_EXACT) };
        }
        Ok(Ctx(p))
    })
    .as_ref()
    .map(|c| c.0)
    .map_err(|e| e.clone())
}

/// The reference puts no upper bound on `w` / `h` (src/lib.rs:61-63), so a request for w = 4e9 reaches
/// this crate.  The library's own bound (IKC_MAX_DIM / IKC_MAX_PIXELS) is checked here, BEFORE the
/// output vector is allocated: the request then fails with a TransformError (HTTP 400) instead of
/// aborting the process in `vec!`.
fn check_dims(sw: u32, sh: u32, dw: u32, dh: u32) -> Result<(), String> {
    let rc = unsafe { ffi::ikc_check_dims(sw, sh, dw, dh) };
    if rc == ffi::IKC_OK { Ok(()) } else { Err(last_error()) }
}

/// Bytes of one tight row, in usize (a u32 product would wrap for wide 16-bit rasters).
fn row_bytes(width: u32, ch: u32, bytes_per_sample: usize) -> usize {
    width as usize * ch as usize * bytes_per_sample
}

fn resize_u8(raw: &[u8], sw: u32, sh: u32, ch: u32, dw: u32, dh: u32) -> Result<Vec<u8>, String> {
    check_dims(sw, sh, dw, dh)?;
    let mut out = vec![0u8; row_bytes(dw, ch, 1) * dh as usize];
    // ikc_submit_u8, not ikc_resize_u8: the handlers (src/lib.rs:180, :286) call this once per request from many tokio
    // workers; the library coalesces the calls that arrive while the GPU is busy into shared uploads and launches.
    let rc = unsafe {
        ffi::ikc_submit_u8(ctx()?, raw.as_ptr(), sw, sh, row_bytes(sw, ch, 1), ch as i32, out.as_mut_ptr(), dw, dh,
                           row_bytes(dw, ch, 1), ffi::IKC_FILTER_LANCZOS3)
    };
    if rc == ffi::IKC_OK { Ok(out) } else { Err(last_error()) }
}

/// Counters of the resize step for imagekit's `/metrics` handler (src/lib.rs:318-338 renders the Prometheus text):
/// images, failures, kernel launches per family, raster bytes, busy time, weight-table hits / misses, coalescing.
pub fn stats() -> Result<ffi::ikc_stats_t, String> {
    let mut s = ffi::ikc_stats_t::default();
    let rc = unsafe { ffi::ikc_get_stats(ctx()?, &mut s) };
    if rc == ffi::IKC_OK { Ok(s) } else { Err(last_error()) }
}

fn resize_u16(raw: &[u16], sw: u32, sh: u32, ch: u32, dw: u32, dh: u32) -> Result<Vec<u16>, String> {
    check_dims(sw, sh, dw, dh)?;
    let mut out = vec![0u16; dw as usize * ch as usize * dh as usize];
    let rc = unsafe {
        ffi::ikc_resize_u16(ctx()?, raw.as_ptr(), sw, sh, row_bytes(sw, ch, 2), ch as i32, out.as_mut_ptr(), dw, dh,
                            row_bytes(dw, ch, 2), ffi::IKC_FILTER_LANCZOS3)
    };
    if rc == ffi::IKC_OK { Ok(out) } else { Err(last_error()) }
}

/// Same contract as imagekit's `resize_image` (src/transform.rs:62-90): `(None, None)` returns the
/// image untouched; otherwise the target size follows the reference's f32 rule, is fitted within
/// (aspect preserved, `DynamicImage::resize`) and the raster is resampled with Lanczos3.
/// The error string becomes `ImageKitError::TransformError` at the call site.
pub fn resize_image(img: DynamicImage, w: Option<u32>, h: Option<u32>) -> Result<DynamicImage, String> {
    if w.is_none() && h.is_none() {
        return Ok(img);
    }
    let (ow, oh) = img.dimensions();
    let (mut tw, mut th) = (0u32, 0u32);
    let code = unsafe {
        ffi::ikc_target_dims(ow, oh, w.is_some() as i32, w.unwrap_or(0), h.is_some() as i32, h.unwrap_or(0), &mut tw, &mut th)
    };
    if code < 0 {
        return Err(last_error());
    }
    if code != ffi::IKC_DIMS_RESAMPLE {
        return Ok(img); // clone / copy cases: pixels unchanged
    }
    macro_rules! run8 {
        ($buf:expr, $ch:expr, $variant:path) => {{
            let v = resize_u8($buf.as_raw(), ow, oh, $ch, tw, th)?;
            Ok($variant(ImageBuffer::from_raw(tw, th, v).expect("size matches")))
        }};
    }
    macro_rules! run16 {
        ($buf:expr, $ch:expr, $variant:path) => {{
            let v = resize_u16($buf.as_raw(), ow, oh, $ch, tw, th)?;
            Ok($variant(ImageBuffer::from_raw(tw, th, v).expect("size matches")))
        }};
    }
    match &img {
        DynamicImage::ImageLuma8(b) => run8!(b, 1, DynamicImage::ImageLuma8),
        DynamicImage::ImageLumaA8(b) => run8!(b, 2, DynamicImage::ImageLumaA8),
        DynamicImage::ImageRgb8(b) => run8!(b, 3, DynamicImage::ImageRgb8),
        DynamicImage::ImageRgba8(b) => run8!(b, 4, DynamicImage::ImageRgba8),
        DynamicImage::ImageLuma16(b) => run16!(b, 1, DynamicImage::ImageLuma16),
        DynamicImage::ImageLumaA16(b) => run16!(b, 2, DynamicImage::ImageLumaA16),
        DynamicImage::ImageRgb16(b) => run16!(b, 3, DynamicImage::ImageRgb16),
        DynamicImage::ImageRgba16(b) => run16!(b, 4, DynamicImage::ImageRgba16),
        // 32F variants never come out of the reference's decoders (jpeg/png/webp); no CPU fallback.
        _ => Err("unsupported pixel format (f32 rasters)".to_string()),
    }
}

/// `resize_image` for a caller that is about to encode the result (the `/img` and `/upload` handlers,
/// src/lib.rs:180-186 and :286-292): `rgba = false` for jpeg/webp, `true` for avif.  The resize stores the
/// variant `encode_image` converts to anyway (`to_rgb8()` / `to_rgba8()`, src/transform.rs:123,131,140), so
/// that conversion becomes a no-op and an RGBA source sends 25 % fewer bytes back from the GPU.
/// 16-bit rasters and the no-resample cases fall through to `resize_image`.
pub fn resize_image_for(img: DynamicImage, w: Option<u32>, h: Option<u32>, rgba: bool) -> Result<DynamicImage, String> {
    let (ow, oh) = img.dimensions();
    let (mut tw, mut th) = (0u32, 0u32);
    let code = unsafe {
        ffi::ikc_target_dims(ow, oh, w.is_some() as i32, w.unwrap_or(0), h.is_some() as i32, h.unwrap_or(0), &mut tw, &mut th)
    };
    let (raw, ch): (&[u8], u32) = match &img {
        DynamicImage::ImageLuma8(b) => (b.as_raw(), 1),
        DynamicImage::ImageLumaA8(b) => (b.as_raw(), 2),
        DynamicImage::ImageRgb8(b) => (b.as_raw(), 3),
        DynamicImage::ImageRgba8(b) => (b.as_raw(), 4),
        _ => return resize_image(img, w, h),
    };
    let co: u32 = if rgba { 4 } else { 3 };
    if (w.is_none() && h.is_none()) || code != ffi::IKC_DIMS_RESAMPLE || co == ch {
        return resize_image(img, w, h);
    }
    check_dims(ow, oh, tw, th)?;  // before the allocation: see check_dims
    let mut out = vec![0u8; row_bytes(tw, co, 1) * th as usize];
    let rc = unsafe {
        ffi::ikc_resize_convert_u8(ctx()?, raw.as_ptr(), ow, oh, row_bytes(ow, ch, 1), ch as i32, out.as_mut_ptr(), tw, th,
                                   row_bytes(tw, co, 1), co as i32, ffi::IKC_FILTER_LANCZOS3)
    };
    if rc != ffi::IKC_OK {
        return Err(last_error());
    }
    Ok(if rgba {
        DynamicImage::ImageRgba8(ImageBuffer::from_raw(tw, th, out).expect("size matches"))
    } else {
        DynamicImage::ImageRgb8(ImageBuffer::from_raw(tw, th, out).expect("size matches"))
    })
}

/// Batched `resize_image`: N independent uploads in one call, sharded round-robin over the box's GPUs
/// (job i -> device i mod G) by `ikc_resize_batch`; no collective, images never leave their GPU.
/// 8-bit variants only; an entry is `Err` if its image could not be resized (the others still are).
/// `(None, None)` entries and same-size targets come back unchanged, as in `resize_image`.
pub fn resize_batch(images: Vec<DynamicImage>, targets: &[(Option<u32>, Option<u32>)]) -> Vec<Result<DynamicImage, String>> {
    assert_eq!(images.len(), targets.len());
    let ctx = match ctx() {
        Ok(c) => c,
        Err(e) => return images.iter().map(|_| Err(e.clone())).collect(),
    };
    struct Slot { job: Option<usize>, ch: u32, tw: u32, th: u32, out: Vec<u8>, err: Option<String> }
    let mut slots: Vec<Slot> = Vec::with_capacity(images.len());
    let mut jobs: Vec<ffi::ikc_job> = Vec::new();
    for (img, &(w, h)) in images.iter().zip(targets) {
        let (ow, oh) = img.dimensions();
        let (mut tw, mut th) = (0u32, 0u32);
        let code = unsafe {
            ffi::ikc_target_dims(ow, oh, w.is_some() as i32, w.unwrap_or(0), h.is_some() as i32, h.unwrap_or(0), &mut tw, &mut th)
        };
        let raw: Option<(&[u8], u32)> = match img {
            DynamicImage::ImageLuma8(b) => Some((b.as_raw(), 1)),
            DynamicImage::ImageLumaA8(b) => Some((b.as_raw(), 2)),
            DynamicImage::ImageRgb8(b) => Some((b.as_raw(), 3)),
            DynamicImage::ImageRgba8(b) => Some((b.as_raw(), 4)),
            _ => None,
        };
        match raw {
            Some((bytes, ch)) if (w.is_some() || h.is_some()) && code == ffi::IKC_DIMS_RESAMPLE => {
                if let Err(e) = check_dims(ow, oh, tw, th) {  // before the allocation: see check_dims
                    slots.push(Slot { job: None, ch, tw, th, out: Vec::new(), err: Some(e) });
                    continue;
                }
                let mut out = vec![0u8; row_bytes(tw, ch, 1) * th as usize];
                jobs.push(ffi::ikc_job {
                    src: bytes.as_ptr() as *const _, dst: out.as_mut_ptr() as *mut _, sw: ow, sh: oh, dw: tw, dh: th,
                    src_pitch: row_bytes(ow, ch, 1), dst_pitch: row_bytes(tw, ch, 1), channels: ch as i32,
                    filter: ffi::IKC_FILTER_LANCZOS3, status: 0, device: 0,
                });
                slots.push(Slot { job: Some(jobs.len() - 1), ch, tw, th, out, err: None });
            }
            _ => slots.push(Slot { job: None, ch: 0, tw, th, out: Vec::new(), err: None }),
        }
    }
    if !jobs.is_empty() {
        unsafe { ffi::ikc_resize_batch(ctx, jobs.as_mut_ptr(), jobs.len()) };  // per-job results are in jobs[i].status
    }
    images.into_iter().zip(targets).zip(slots).map(|((img, &(w, h)), s)| match s.job {
        None if s.err.is_some() => Err(s.err.unwrap()),
        None => resize_image(img, w, h),  // passthrough, clone, 16-bit: the single-image path
        Some(j) if jobs[j].status != ffi::IKC_OK => Err(format!("resize failed with status {}", jobs[j].status)),
        Some(_) => {
            let out = s.out;
            Ok(match s.ch {
                1 => DynamicImage::ImageLuma8(ImageBuffer::from_raw(s.tw, s.th, out).expect("size matches")),
                2 => DynamicImage::ImageLumaA8(ImageBuffer::from_raw(s.tw, s.th, out).expect("size matches")),
                3 => DynamicImage::ImageRgb8(ImageBuffer::from_raw(s.tw, s.th, out).expect("size matches")),
                _ => DynamicImage::ImageRgba8(ImageBuffer::from_raw(s.tw, s.th, out).expect("size matches")),
            })
        }
    }).collect()
}

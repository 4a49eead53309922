//! Reference outputs for the golden inputs, from the `image` crate itself (0.25.8, the version the reference locks).
//!
//! Manifest line:  <name> <h> <w> <channels> <dw> <dh> <filter 0..4> <mode>
//!   mode "exact": image::imageops::resize(&buf, dw, dh, filter)           -- imageops/sample.rs, what
//!                 DynamicImage::resize_exact runs (reference src/transform.rs:85-89 reaches it through resize)
//!   mode "fit"  : DynamicImage::resize(dw, dh, filter)                    -- the call the reference makes: fit within
//!                 (dw, dh) keeping the aspect ratio, then the same resampling; the output size is written to
//!                 <name>.ref.dims as "w h"
//! Input  <dir>/<name>.src.bin : h * w * channels bytes, row-major, interleaved.
//! Output <dir>/<name>.ref.bin : the resampled raster, same layout.
use image::imageops::{self, FilterType};
use image::{DynamicImage, ImageBuffer, Luma, LumaA, Pixel, Rgb, Rgba};
use std::{env, fs, path::Path};

fn filter(code: u32) -> FilterType {
    match code {
        0 => FilterType::Nearest,
        1 => FilterType::Triangle,
        2 => FilterType::CatmullRom,
        3 => FilterType::Gaussian,
        _ => FilterType::Lanczos3,
    }
}

fn exact<P>(raw: Vec<u8>, w: u32, h: u32, dw: u32, dh: u32, f: FilterType) -> Vec<u8>
where
    P: Pixel<Subpixel = u8> + 'static,
{
    let img: ImageBuffer<P, Vec<u8>> = ImageBuffer::from_raw(w, h, raw).expect("raster size does not match the manifest");
    imageops::resize(&img, dw, dh, f).into_raw()
}

fn fit(raw: Vec<u8>, w: u32, h: u32, c: u32, dw: u32, dh: u32, f: FilterType) -> (Vec<u8>, u32, u32) {
    let img = match c {
        1 => DynamicImage::ImageLuma8(ImageBuffer::from_raw(w, h, raw).expect("size")),
        2 => DynamicImage::ImageLumaA8(ImageBuffer::from_raw(w, h, raw).expect("size")),
        3 => DynamicImage::ImageRgb8(ImageBuffer::from_raw(w, h, raw).expect("size")),
        _ => DynamicImage::ImageRgba8(ImageBuffer::from_raw(w, h, raw).expect("size")),
    };
    let out = img.resize(dw, dh, f);
    let (ow, oh) = (out.width(), out.height());
    (out.into_bytes(), ow, oh)
}

fn main() {
    let dir = env::args().nth(1).unwrap_or_else(|| "inputs".to_string());
    let dir = Path::new(&dir);
    let manifest = fs::read_to_string(dir.join("cases.txt")).expect("cases.txt (run tests/golden/export_inputs.py first)");
    let mut n = 0;
    for line in manifest.lines() {
        let t: Vec<&str> = line.split_whitespace().collect();
        if t.len() != 8 || t[0].starts_with('#') {
            continue;
        }
        let name = t[0];
        let num = |i: usize| t[i].parse::<u32>().expect("number");
        let (h, w, c, dw, dh, f) = (num(1), num(2), num(3), num(4), num(5), filter(num(6)));
        let raw = fs::read(dir.join(format!("{name}.src.bin"))).expect("input raster");
        let out = if t[7] == "fit" {
            let (bytes, ow, oh) = fit(raw, w, h, c, dw, dh, f);
            fs::write(dir.join(format!("{name}.ref.dims")), format!("{ow} {oh}\n")).expect("write dims");
            bytes
        } else {
            match c {
                1 => exact::<Luma<u8>>(raw, w, h, dw, dh, f),
                2 => exact::<LumaA<u8>>(raw, w, h, dw, dh, f),
                3 => exact::<Rgb<u8>>(raw, w, h, dw, dh, f),
                _ => exact::<Rgba<u8>>(raw, w, h, dw, dh, f),
            }
        };
        fs::write(dir.join(format!("{name}.ref.bin")), out).expect("write output");
        n += 1;
    }
    println!("wrote {n} reference rasters with image {}", "0.25.8");
}

//! Raw bindings of include/imagekit_cuda.h (the C ABI of libimagekit_cuda.so).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct ikc_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct ikc_batch {
    _private: [u8; 0],
}

pub const IKC_OK: c_int = 0;
pub const IKC_FILTER_LANCZOS3: c_int = 4;
pub const IKC_DIMS_RESAMPLE: c_int = 0;
pub const IKC_MODE_FAST: c_int = 0;
pub const IKC_MODE_EXACT: c_int = 1;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ikc_job {
    pub src: *const c_void,
    pub dst: *mut c_void,
    pub sw: u32,
    pub sh: u32,
    pub dw: u32,
    pub dh: u32,
    pub src_pitch: usize,
    pub dst_pitch: usize,
    pub channels: i32,
    pub filter: i32,
    pub status: i32,
    pub device: i32,
}

/// struct ikc_stats_t: monotonic counters for a /metrics handler.
#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct ikc_stats_t {
    pub calls: u64,
    pub failed: u64,
    pub trivial: u64,
    pub launches: u64,
    pub launches_banded8t: u64,
    pub launches_banded8: u64,
    pub launches_banded_f16: u64,
    pub launches_ring: u64,
    pub launches_up2: u64,
    pub launches_tile: u64,
    pub launches_generic: u64,
    pub src_bytes: u64,
    pub dst_bytes: u64,
    pub busy_ns: u64,
    pub table_hits: u64,
    pub table_misses: u64,
    pub submit_batches: u64,
    pub submit_jobs: u64,
    pub launches_banded8u: u64,
    pub staging_trims: u64,
}

extern "C" {
    pub fn ikc_create(device_ids: *const c_int, n: c_int, out: *mut *mut ikc_ctx) -> c_int;
    pub fn ikc_destroy(ctx: *mut ikc_ctx);
    pub fn ikc_device_count(ctx: *const ikc_ctx) -> c_int;
    pub fn ikc_set_mode(ctx: *mut ikc_ctx, mode: c_int) -> c_int;
    pub fn ikc_last_error() -> *const c_char;
    pub fn ikc_check_dims(sw: u32, sh: u32, dw: u32, dh: u32) -> c_int;
    pub fn ikc_target_dims(ow: u32, oh: u32, has_w: c_int, w: u32, has_h: c_int, h: u32, tw: *mut u32, th: *mut u32) -> c_int;
    pub fn ikc_resize_u8(ctx: *mut ikc_ctx, src: *const u8, sw: u32, sh: u32, src_pitch: usize, channels: c_int,
                         dst: *mut u8, dw: u32, dh: u32, dst_pitch: usize, filter: c_int) -> c_int;
    pub fn ikc_submit_u8(ctx: *mut ikc_ctx, src: *const u8, sw: u32, sh: u32, src_pitch: usize, channels: c_int,
                         dst: *mut u8, dw: u32, dh: u32, dst_pitch: usize, filter: c_int) -> c_int;
    pub fn ikc_get_stats(ctx: *const ikc_ctx, out: *mut ikc_stats_t) -> c_int;
    pub fn ikc_resize_convert_u8(ctx: *mut ikc_ctx, src: *const u8, sw: u32, sh: u32, src_pitch: usize, src_channels: c_int,
                                 dst: *mut u8, dw: u32, dh: u32, dst_pitch: usize, dst_channels: c_int, filter: c_int) -> c_int;
    pub fn ikc_resize_u16(ctx: *mut ikc_ctx, src: *const u16, sw: u32, sh: u32, src_pitch: usize, channels: c_int,
                          dst: *mut u16, dw: u32, dh: u32, dst_pitch: usize, filter: c_int) -> c_int;
    pub fn ikc_resize_batch(ctx: *mut ikc_ctx, jobs: *mut ikc_job, n: usize) -> c_int;
    pub fn ikc_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn ikc_host_free(p: *mut c_void);
}

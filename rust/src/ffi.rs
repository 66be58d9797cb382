//! Raw bindings of include/phastft_cuda.h (hand-written; one line per C declaration).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct phastft_plan_dit_f64 { _p: [u8; 0] }
#[repr(C)] pub struct phastft_plan_dit_f32 { _p: [u8; 0] }
#[repr(C)] pub struct phastft_plan_r2c_f64 { _p: [u8; 0] }
#[repr(C)] pub struct phastft_plan_r2c_f32 { _p: [u8; 0] }

#[repr(C)]
pub struct phastft_options { pub multithreaded_bit_reversal: i32, pub smallest_parallel_chunk_size: usize }

extern "C" {
    pub fn phastft_last_error() -> *const c_char;
    pub fn phastft_options_guess(input_size: usize, out: *mut phastft_options);
    pub fn phastft_oneshot_cache_clear();
    pub fn phastft_host_register(host_ptr: *mut c_void, bytes: usize) -> i32;
    pub fn phastft_host_unregister(host_ptr: *mut c_void) -> i32;
    pub fn phastft_fft_dit_f64_oneshot(re: *mut f64, len_re: usize, im: *mut f64, len_im: usize, direction: c_int, device: c_int) -> i32;
    pub fn phastft_fft_dit_f32_oneshot(re: *mut f32, len_re: usize, im: *mut f32, len_im: usize, direction: c_int, device: c_int) -> i32;

    pub fn phastft_plan_dit_f64_create(n: usize, device: c_int, mode: c_int, out: *mut *mut phastft_plan_dit_f64) -> i32;
    pub fn phastft_plan_dit_f32_create(n: usize, device: c_int, mode: c_int, out: *mut *mut phastft_plan_dit_f32) -> i32;
    pub fn phastft_plan_dit_f64_destroy(p: *mut phastft_plan_dit_f64);
    pub fn phastft_plan_dit_f32_destroy(p: *mut phastft_plan_dit_f32);
    pub fn phastft_fft_dit_f64_host(p: *const phastft_plan_dit_f64, re: *mut f64, len_re: usize, im: *mut f64, len_im: usize,
                                    direction: c_int, opts: *const phastft_options) -> i32;
    pub fn phastft_fft_dit_f32_host(p: *const phastft_plan_dit_f32, re: *mut f32, len_re: usize, im: *mut f32, len_im: usize,
                                    direction: c_int, opts: *const phastft_options) -> i32;
    pub fn phastft_fft_dit_f64_dev(p: *const phastft_plan_dit_f64, d_re: *mut f64, d_im: *mut f64, direction: c_int,
                                   batch: usize, batch_stride: usize, stream: *mut c_void) -> i32;
    pub fn phastft_fft_dit_f32_dev(p: *const phastft_plan_dit_f32, d_re: *mut f32, d_im: *mut f32, direction: c_int,
                                   batch: usize, batch_stride: usize, stream: *mut c_void) -> i32;
    pub fn phastft_fft_interleaved_f64_host(p: *const phastft_plan_dit_f64, signal: *mut f64, len_complex: usize, direction: c_int) -> i32;
    pub fn phastft_fft_interleaved_f32_host(p: *const phastft_plan_dit_f32, signal: *mut f32, len_complex: usize, direction: c_int) -> i32;

    pub fn phastft_plan_r2c_f64_create(n: usize, device: c_int, out: *mut *mut phastft_plan_r2c_f64) -> i32;
    pub fn phastft_plan_r2c_f32_create(n: usize, device: c_int, out: *mut *mut phastft_plan_r2c_f32) -> i32;
    pub fn phastft_plan_r2c_f64_destroy(p: *mut phastft_plan_r2c_f64);
    pub fn phastft_plan_r2c_f32_destroy(p: *mut phastft_plan_r2c_f32);
    pub fn phastft_r2c_f64_host(p: *const phastft_plan_r2c_f64, input: *const f64, len_in: usize, out_re: *mut f64, len_ore: usize,
                                out_im: *mut f64, len_oim: usize) -> i32;
    pub fn phastft_r2c_f32_host(p: *const phastft_plan_r2c_f32, input: *const f32, len_in: usize, out_re: *mut f32, len_ore: usize,
                                out_im: *mut f32, len_oim: usize) -> i32;
    pub fn phastft_c2r_f64_host(p: *const phastft_plan_r2c_f64, in_re: *const f64, len_ire: usize, in_im: *const f64, len_iim: usize,
                                output: *mut f64, len_out: usize, scratch_re: *mut f64, len_sre: usize, scratch_im: *mut f64, len_sim: usize) -> i32;
    pub fn phastft_c2r_f32_host(p: *const phastft_plan_r2c_f32, in_re: *const f32, len_ire: usize, in_im: *const f32, len_iim: usize,
                                output: *mut f32, len_out: usize, scratch_re: *mut f32, len_sre: usize, scratch_im: *mut f32, len_sim: usize) -> i32;
}

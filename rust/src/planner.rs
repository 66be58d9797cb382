//! `planner.rs` of the reference: `Direction`, `PlannerMode`, `PlannerDit{64,32}`, `PlannerR2c{64,32}`.
//! The planners own an opaque handle to the CUDA plan (device twiddle tables, workspace, stream) and free
//! it on drop.  They stay `Send + Sync` like the reference's (plain `Vec` fields there): the C library
//! serialises workspace use inside the plan.
use crate::{check, device, ffi};

/// `planner.rs:10-16`
#[derive(Copy, Clone)]
pub enum Direction {
    Forward = 1,
    Reverse = -1,
}

/// `planner.rs:25-32` (accepted and ignored, as in the reference: `planner.rs:65`)
#[derive(Copy, Clone, Debug, Default)]
pub enum PlannerMode {
    #[default]
    Heuristic,
    Tune,
}

macro_rules! impl_planner_dit {
    ($name:ident, $raw:ty, $create:ident, $destroy:ident) => {
        pub struct $name {
            pub(crate) raw: *mut $raw,
        }
        unsafe impl Send for $name {}
        unsafe impl Sync for $name {}
        impl $name {
            /// `planner.rs:55`: panics unless `num_points` is a non-zero power of two.
            pub fn new(num_points: usize) -> Self {
                Self::with_mode(num_points, PlannerMode::Heuristic)
            }
            /// `planner.rs:65`
            pub fn with_mode(num_points: usize, mode: PlannerMode) -> Self {
                let mut raw = std::ptr::null_mut();
                check(unsafe { ffi::$create(num_points, device(), mode as i32, &mut raw) });
                Self { raw }
            }
        }
        impl Drop for $name {
            fn drop(&mut self) {
                unsafe { ffi::$destroy(self.raw) }
            }
        }
    };
}
impl_planner_dit!(PlannerDit64, ffi::phastft_plan_dit_f64, phastft_plan_dit_f64_create, phastft_plan_dit_f64_destroy);
impl_planner_dit!(PlannerDit32, ffi::phastft_plan_dit_f32, phastft_plan_dit_f32_create, phastft_plan_dit_f32_destroy);

macro_rules! impl_planner_r2c {
    ($name:ident, $raw:ty, $create:ident, $destroy:ident) => {
        pub struct $name {
            pub(crate) raw: *mut $raw,
        }
        unsafe impl Send for $name {}
        unsafe impl Sync for $name {}
        impl $name {
            /// `planner.rs:194`: panics with "n must be a power of 2 >= 4".
            pub fn new(n: usize) -> Self {
                let mut raw = std::ptr::null_mut();
                check(unsafe { ffi::$create(n, device(), &mut raw) });
                Self { raw }
            }
        }
        impl Drop for $name {
            fn drop(&mut self) {
                unsafe { ffi::$destroy(self.raw) }
            }
        }
    };
}
impl_planner_r2c!(PlannerR2c64, ffi::phastft_plan_r2c_f64, phastft_plan_r2c_f64_create, phastft_plan_r2c_f64_destroy);
impl_planner_r2c!(PlannerR2c32, ffi::phastft_plan_r2c_f32, phastft_plan_r2c_f32_create, phastft_plan_r2c_f32_destroy);

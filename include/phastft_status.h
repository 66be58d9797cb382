/*
 * phastft_status.h -- status codes shared by every entry point of libphastft_cuda.so.
 *
 * The reference (QuState/PhastFT, Rust) signals misuse with `assert!`/`assert_eq!`
 * panics.  A C ABI cannot unwind, so every entry point returns one of these codes
 * and the host-language wrapper (rust/src/lib.rs, cpp/phastft.hpp,
 * phastft_b200/api.py) turns a non-zero code back into a panic / exception that
 * carries the reference's exact message (phastft_status_message()).
 *
 * One code per reference panic site:
 *   src/algorithms/dit.rs:284      assert_eq!(reals.len(), imags.len())
 *   src/algorithms/dit.rs:285      assert!(reals.len().is_power_of_two())
 *   src/planner.rs:66              assert!(num_points > 0 && num_points.is_power_of_two())
 *   src/algorithms/dit.rs:289      assert_eq!(log_n, planner.log_n)
 *   src/planner.rs:195             "n must be a power of 2 >= 4"
 *   src/algorithms/r2c.rs:543-553  r2c length asserts
 *   src/algorithms/r2c.rs:750-762  c2r length asserts
 */
#ifndef PHASTFT_STATUS_H
#define PHASTFT_STATUS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum phastft_status {
    PHASTFT_OK = 0,
    PHASTFT_ERR_LEN_MISMATCH = 1,   /* reals.len() != imags.len()              (dit.rs:284)  */
    PHASTFT_ERR_NOT_POW2 = 2,       /* length / num_points not a power of two  (dit.rs:285, planner.rs:66) */
    PHASTFT_ERR_PLAN_MISMATCH = 3,  /* log2(len) != planner.log_n              (dit.rs:289)  */
    PHASTFT_ERR_R2C_N = 4,          /* "n must be a power of 2 >= 4"           (planner.rs:195) */
    PHASTFT_ERR_INPUT_LEN = 5,      /* "input length must match planner size"  (r2c.rs:543)  */
    PHASTFT_ERR_OUTPUT_RE_LEN = 6,  /* "output_re must have length N/2 + 1"    (r2c.rs:544)  */
    PHASTFT_ERR_OUTPUT_IM_LEN = 7,  /* "output_im must have length N/2 + 1"    (r2c.rs:549)  */
    PHASTFT_ERR_OUTPUT_LEN = 8,     /* "output length must match planner size" (r2c.rs:750)  */
    PHASTFT_ERR_INPUT_RE_LEN = 9,   /* "input_re must have length N/2 + 1"     (r2c.rs:751)  */
    PHASTFT_ERR_INPUT_IM_LEN = 10,  /* "input_im must have length N/2 + 1"     (r2c.rs:756)  */
    PHASTFT_ERR_SCRATCH_RE_LEN = 11,/* "scratch_re must have length N/2"       (r2c.rs:761)  */
    PHASTFT_ERR_SCRATCH_IM_LEN = 12,/* "scratch_im must have length N/2"       (r2c.rs:762)  */
    PHASTFT_ERR_INVALID_ARG = 13,   /* NULL pointer, bad direction, bad batch stride (no reference analogue: Rust's type system rules these out) */
    PHASTFT_ERR_CUDA = 100,         /* a CUDA runtime call failed; see phastft_last_error() */
    PHASTFT_ERR_NCCL = 101,         /* an NCCL call failed */
    PHASTFT_ERR_NO_DEVICE = 102     /* no CUDA device / driver: the library never falls back to the CPU */
} phastft_status;

/* Direction discriminants are the reference's: planner.rs:11-16 */
#define PHASTFT_FORWARD 1
#define PHASTFT_REVERSE (-1)

/* The reference's panic message for a status code ("" for PHASTFT_OK).  Static storage. */
static inline const char* phastft_status_message(int32_t code) {
    switch (code) {
        case PHASTFT_OK: return "";
        case PHASTFT_ERR_LEN_MISMATCH: return "assertion `left == right` failed: reals.len() == imags.len()";
        case PHASTFT_ERR_NOT_POW2: return "assertion failed: length must be a non-zero power of two";
        case PHASTFT_ERR_PLAN_MISMATCH: return "assertion `left == right` failed: log_n == planner.log_n";
        case PHASTFT_ERR_R2C_N: return "n must be a power of 2 >= 4";
        case PHASTFT_ERR_INPUT_LEN: return "input length must match planner size";
        case PHASTFT_ERR_OUTPUT_RE_LEN: return "output_re must have length N/2 + 1";
        case PHASTFT_ERR_OUTPUT_IM_LEN: return "output_im must have length N/2 + 1";
        case PHASTFT_ERR_OUTPUT_LEN: return "output length must match planner size";
        case PHASTFT_ERR_INPUT_RE_LEN: return "input_re must have length N/2 + 1";
        case PHASTFT_ERR_INPUT_IM_LEN: return "input_im must have length N/2 + 1";
        case PHASTFT_ERR_SCRATCH_RE_LEN: return "scratch_re must have length N/2";
        case PHASTFT_ERR_SCRATCH_IM_LEN: return "scratch_im must have length N/2";
        case PHASTFT_ERR_INVALID_ARG: return "invalid argument";
        case PHASTFT_ERR_CUDA: return "CUDA error";
        case PHASTFT_ERR_NCCL: return "NCCL error";
        case PHASTFT_ERR_NO_DEVICE: return "no CUDA device available (phastft_cuda has no CPU fallback)";
        default: return "unknown phastft status";
    }
}

#ifdef __cplusplus
}
#endif
#endif /* PHASTFT_STATUS_H */

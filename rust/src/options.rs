//! `options.rs:10-43` of the reference, unchanged in shape.  On the GPU both fields are hints with no
//! effect: there is no separate bit-reversal step to thread, and every launch is full-chip parallel.

#[non_exhaustive]
#[derive(Debug, Clone)]
pub struct Options {
    pub multithreaded_bit_reversal: bool,
    pub smallest_parallel_chunk_size: usize,
}

impl Default for Options {
    fn default() -> Self {
        Self { multithreaded_bit_reversal: false, smallest_parallel_chunk_size: 16384 }
    }
}

impl Options {
    /// `options.rs:38-43`
    pub fn guess_options(input_size: usize) -> Options {
        let mut o = Options::default();
        if input_size > 0 {
            o.multithreaded_bit_reversal = input_size.ilog2() as usize >= 16;
        }
        o
    }
}

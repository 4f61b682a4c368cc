//! `resampler-cuda`: the reference's `ResamplerFir` API on a B200.
//!
//! Source only in this repository (no rustc in the build image; see `build.rs`).  It shows the
//! reference-side binding of every entry point of `include/resampler_b200.h`.
//!
//! ```ignore
//! use resampler_cuda::{Attenuation, Latency, ResamplerFir, SampleRate};
//! let mut r = ResamplerFir::new(2, SampleRate::Hz44100, SampleRate::Hz48000,
//!                               Latency::Sample64, Attenuation::Db90);
//! let mut out = vec![0.0f32; r.buffer_size_output()];
//! let (consumed, produced) = r.resample(&input, &mut out)?;
//! ```
use std::os::raw::{c_int, c_void};

pub use resampler::{Attenuation, Latency, ResampleError, SampleRate};

#[repr(C)]
struct RsbFir {
    _private: [u8; 0],
}

const RSB_MEM_DEVICE: c_int = 0;
const RSB_MEM_HOST: c_int = 1;

unsafe extern "C" {
    fn rsb_fir_create(out: *mut *mut RsbFir, device: c_int, n_streams: u32, channels: u32,
                      input_rate_hz: u32, output_rate_hz: u32, latency: c_int, attenuation: c_int) -> c_int;
    fn rsb_fir_destroy(h: *mut RsbFir);
    fn rsb_fir_buffer_size_output(h: *const RsbFir) -> usize;
    fn rsb_fir_delay(h: *const RsbFir) -> usize;
    fn rsb_fir_reset(h: *mut RsbFir, stream: i64) -> c_int;
    fn rsb_fir_resample(h: *mut RsbFir, stream: u32, input: *const f32, input_len: usize,
                        output: *mut f32, output_len: usize, consumed: *mut usize,
                        produced: *mut usize) -> c_int;
    fn rsb_fir_submit_batch(h: *mut RsbFir, n: u32, streams: *const u32, inp: *const *const f32,
                            in_lens: *const usize, out: *const *mut f32, out_lens: *const usize,
                            consumed: *mut usize, produced: *mut usize, memspace: c_int, flags: u32) -> c_int;
    fn rsb_fir_process_batch(h: *mut RsbFir, n: u32, streams: *const u32, inp: *const *const f32,
                             total_lens: *const usize, call_len: usize, out_cap_len: usize,
                             out: *const *mut f32, out_capacities: *const usize,
                             consumed_totals: *mut usize, produced_totals: *mut usize,
                             n_calls: *mut u32, memspace: c_int, flags: u32) -> c_int;
    fn rsb_fir_process_pcm_batch(h: *mut RsbFir, n: u32, streams: *const u32, inp: *const *const c_void,
                                 in_frames: *const usize, format: c_int, src_channels: u32,
                                 call_len: usize, out_cap_len: usize, out: *const *mut f32,
                                 out_capacities: *const usize, consumed_totals: *mut usize,
                                 produced_totals: *mut usize, n_calls: *mut u32, memspace: c_int,
                                 flags: u32) -> c_int;
    fn rsb_fir_flush_batch(h: *mut RsbFir, n: u32, streams: *const u32, out: *const *mut f32,
                           out_capacities: *const usize, produced: *mut usize, memspace: c_int,
                           flags: u32) -> c_int;
    fn rsb_fir_sync(h: *mut RsbFir) -> c_int;
    fn rsb_alloc_pinned(bytes: usize) -> *mut c_void;
    fn rsb_free_pinned(p: *mut c_void);
    fn rsb_status_string(status: c_int) -> *const std::os::raw::c_char;
}

fn latency_code(l: Latency) -> c_int {
    match l { Latency::Sample8 => 0, Latency::Sample16 => 1, Latency::Sample32 => 2, Latency::Sample64 => 3 }
}
fn attenuation_code(a: Attenuation) -> c_int {
    match a { Attenuation::Db60 => 0, Attenuation::Db90 => 1, Attenuation::Db120 => 2 }
}
fn map_err(rc: c_int) -> Result<(), ResampleError> {
    match rc {
        0 => Ok(()),
        1 => Err(ResampleError::InvalidInputBufferSize),
        2 => Err(ResampleError::InvalidOutputBufferSize),
        // CUDA / device failures have no counterpart in the reference's error type
        other => panic!("resampler-cuda: device error {other}"),
    }
}

/// N independent streams with identical parameters on one GPU.
pub struct FirBatch {
    h: *mut RsbFir,
    channels: usize,
}
// one caller at a time per handle (`&mut self`), handles may move between threads
unsafe impl Send for FirBatch {}

impl FirBatch {
    pub fn new_from_hz(n_streams: u32, channels: usize, input_rate_hz: u32, output_rate_hz: u32,
                       latency: Latency, attenuation: Attenuation, device: i32) -> Self {
        // same panics, same texts, same order as resampler_fir.rs:302-309
        assert!(input_rate_hz > 0, "input sample rate must be greater than zero");
        assert!(output_rate_hz > 0, "output sample rate must be greater than zero");
        let mut h = std::ptr::null_mut();
        let rc = unsafe {
            rsb_fir_create(&mut h, device, n_streams, channels as u32, input_rate_hz, output_rate_hz,
                           latency_code(latency), attenuation_code(attenuation))
        };
        assert!(rc == 0, "resampler-cuda: create failed ({rc})");
        Self { h, channels }
    }
    pub fn buffer_size_output(&self) -> usize { unsafe { rsb_fir_buffer_size_output(self.h) } }
    pub fn delay(&self) -> usize { unsafe { rsb_fir_delay(self.h) } }
    pub fn reset(&mut self, stream: Option<u32>) {
        unsafe { rsb_fir_reset(self.h, stream.map(|s| s as i64).unwrap_or(-1)) };
    }
    /// One `resample()` call per listed stream (host slices).
    pub fn submit(&mut self, streams: &[u32], inputs: &[&[f32]], outputs: &mut [&mut [f32]])
                  -> Result<Vec<(usize, usize)>, ResampleError> {
        let n = inputs.len();
        let in_ptrs: Vec<*const f32> = inputs.iter().map(|s| s.as_ptr()).collect();
        let in_lens: Vec<usize> = inputs.iter().map(|s| s.len()).collect();
        let out_ptrs: Vec<*mut f32> = outputs.iter_mut().map(|s| s.as_mut_ptr()).collect();
        let out_lens: Vec<usize> = outputs.iter().map(|s| s.len()).collect();
        let (mut c, mut p) = (vec![0usize; n], vec![0usize; n]);
        map_err(unsafe {
            rsb_fir_submit_batch(self.h, n as u32, streams.as_ptr(), in_ptrs.as_ptr(), in_lens.as_ptr(),
                                 out_ptrs.as_ptr(), out_lens.as_ptr(), c.as_mut_ptr(), p.as_mut_ptr(),
                                 RSB_MEM_HOST, 0)
        })?;
        Ok(c.into_iter().zip(p).collect())
    }
    /// The CLI's batch path for many files at once (resample/src/main.rs:128-156 + 226-254):
    /// `inputs[i]` are one file's raw little-endian sample bytes as hound would decode them
    /// (`format`: 0 = u8, 1 = s16, 2 = packed s24, 3 = s32, 4 = f32), `src_channels` 1 (mono,
    /// duplicated into every channel on the GPU) or `channels`.  Output vectors hold the
    /// resampled interleaved f32 signal of each file (512-value calls like the CLI).
    pub fn process_pcm(&mut self, inputs: &[&[u8]], format: i32, src_channels: u32)
                       -> Result<Vec<Vec<f32>>, ResampleError> {
        let n = inputs.len();
        let bps = [1usize, 2, 3, 4, 4][format as usize];
        let frames: Vec<usize> = inputs.iter().map(|b| b.len() / (bps * src_channels as usize)).collect();
        let ratio_bound = self.buffer_size_output() / self.channels;   // frames per 4096-frame call
        let caps: Vec<usize> = frames.iter()
            .map(|f| ((f / 3968 + 2) * ratio_bound + 8) * self.channels).collect();
        let mut outs: Vec<Vec<f32>> = caps.iter().map(|c| vec![0.0f32; *c]).collect();
        let in_ptrs: Vec<*const c_void> = inputs.iter().map(|b| b.as_ptr() as *const c_void).collect();
        let out_ptrs: Vec<*mut f32> = outs.iter_mut().map(|v| v.as_mut_ptr()).collect();
        let mut produced = vec![0usize; n];
        map_err(unsafe {
            rsb_fir_process_pcm_batch(self.h, n as u32, std::ptr::null(), in_ptrs.as_ptr(), frames.as_ptr(),
                                      format, src_channels, 512, 0, out_ptrs.as_ptr(), caps.as_ptr(),
                                      std::ptr::null_mut(), produced.as_mut_ptr(), std::ptr::null_mut(),
                                      RSB_MEM_HOST, 0)
        })?;
        for (v, p) in outs.iter_mut().zip(&produced) { v.truncate(*p); }
        Ok(outs)
    }
    /// Opt-in tail handling (not in the reference): `delay()` frames of silence per stream.
    pub fn flush(&mut self, outputs: &mut [&mut [f32]]) -> Result<Vec<usize>, ResampleError> {
        let n = outputs.len();
        let out_ptrs: Vec<*mut f32> = outputs.iter_mut().map(|s| s.as_mut_ptr()).collect();
        let caps: Vec<usize> = outputs.iter().map(|s| s.len()).collect();
        let mut produced = vec![0usize; n];
        map_err(unsafe {
            rsb_fir_flush_batch(self.h, n as u32, std::ptr::null(), out_ptrs.as_ptr(), caps.as_ptr(),
                                produced.as_mut_ptr(), RSB_MEM_HOST, 0)
        })?;
        Ok(produced)
    }
    pub fn channels(&self) -> usize { self.channels }
}
impl Drop for FirBatch {
    fn drop(&mut self) { unsafe { rsb_fir_destroy(self.h) } }
}

/// Drop-in for `resampler::ResamplerFir` (src/resampler_fir.rs:179-201).
pub struct ResamplerFir { inner: FirBatch }

impl ResamplerFir {
    /// `ResamplerFir::new` (resampler_fir.rs:252-266)
    pub fn new(channels: usize, input_rate: SampleRate, output_rate: SampleRate, latency: Latency,
               attenuation: Attenuation) -> Self {
        Self::new_from_hz(channels, u32::from(input_rate), u32::from(output_rate), latency, attenuation)
    }
    /// `ResamplerFir::new_from_hz` (resampler_fir.rs:295-404)
    pub fn new_from_hz(channels: usize, input_rate_hz: u32, output_rate_hz: u32, latency: Latency,
                       attenuation: Attenuation) -> Self {
        Self { inner: FirBatch::new_from_hz(1, channels, input_rate_hz, output_rate_hz, latency, attenuation, 0) }
    }
    /// resampler_fir.rs:456-465
    pub fn buffer_size_output(&self) -> usize { self.inner.buffer_size_output() }
    /// resampler_fir.rs:509-621
    pub fn resample(&mut self, input: &[f32], output: &mut [f32]) -> Result<(usize, usize), ResampleError> {
        let (mut c, mut p) = (0usize, 0usize);
        map_err(unsafe {
            rsb_fir_resample(self.inner.h, 0, input.as_ptr(), input.len(), output.as_mut_ptr(),
                             output.len(), &mut c, &mut p)
        })?;
        Ok((c, p))
    }
    /// resampler_fir.rs:630-632
    pub fn delay(&self) -> usize { self.inner.delay() }
    /// resampler_fir.rs:638-642
    pub fn reset(&mut self) { self.inner.reset(Some(0)) }
}

#[allow(dead_code)]
fn _unused() {
    // keeps the remaining bindings referenced (device-memory batch entry points, pinned helpers)
    let _ = (rsb_fir_process_batch as usize, rsb_fir_sync as usize, rsb_alloc_pinned as usize,
             rsb_free_pinned as usize, rsb_status_string as usize, RSB_MEM_DEVICE);
}

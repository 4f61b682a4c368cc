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
    fn rsb_fir_cuda_stream(h: *const RsbFir) -> *mut c_void;
    fn rsb_fir_host_pipeline_stats(h: *const RsbFir, batches: *mut u64, slices: *mut u64) -> c_int;
    fn rsb_alloc_pinned(bytes: usize) -> *mut c_void;
    fn rsb_free_pinned(p: *mut c_void);
    fn rsb_alloc_device(device: c_int, bytes: usize) -> *mut c_void;
    fn rsb_free_device(device: c_int, p: *mut c_void);
    fn rsb_memcpy(device: c_int, dst: *mut c_void, src: *const c_void, bytes: usize, kind: c_int) -> c_int;
    fn rsb_status_string(status: c_int) -> *const std::os::raw::c_char;
    fn rsb_last_error() -> *const std::os::raw::c_char;
    fn rsb_set_device_filter_design(enable: c_int) -> c_int;
    fn rsb_fft_create(out: *mut *mut RsbFft, device: c_int, n_streams: u32, channels: u32,
                      input_rate_hz: u32, output_rate_hz: u32) -> c_int;
    fn rsb_fft_destroy(h: *mut RsbFft);
    fn rsb_fft_chunk_size_input(h: *const RsbFft) -> usize;
    fn rsb_fft_chunk_size_output(h: *const RsbFft) -> usize;
    fn rsb_fft_delay(h: *const RsbFft) -> usize;
    fn rsb_fft_resample(h: *mut RsbFft, stream: u32, input: *const f32, input_len: usize,
                        output: *mut f32, output_len: usize) -> c_int;
}

#[repr(C)]
struct RsbFft {
    _private: [u8; 0],
}

/// Tables missing from the process-wide cache are designed on the GPU from now on (bit-identical to
/// the host design); returns the previous setting.
pub fn set_device_filter_design(enable: bool) -> bool {
    unsafe { rsb_set_device_filter_design(enable as c_int) != 0 }
}

const RSB_FLAG_ASYNC: u32 = 1;

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
        other => {
            let (name, detail) = unsafe {
                (std::ffi::CStr::from_ptr(rsb_status_string(other)).to_string_lossy().into_owned(),
                 std::ffi::CStr::from_ptr(rsb_last_error()).to_string_lossy().into_owned())
            };
            panic!("resampler-cuda: {name} ({other}): {detail}")
        }
    }
}

/// Page-locked host memory of `len` f32 values (RAII over `rsb_alloc_pinned`).  Host-memspace calls
/// on pinned buffers copy at the full PCIe rate and let `process_batch` overlap the copies of one
/// time slice with the kernels of another; pageable slices work too, slower.
pub struct PinnedBuffer {
    ptr: *mut f32,
    len: usize,
}
// plain memory owned by the value; no thread affinity
unsafe impl Send for PinnedBuffer {}
unsafe impl Sync for PinnedBuffer {}
impl PinnedBuffer {
    pub fn new(len: usize) -> Option<Self> {
        let ptr = unsafe { rsb_alloc_pinned(len * std::mem::size_of::<f32>()) } as *mut f32;
        if ptr.is_null() { return None; }
        unsafe { std::ptr::write_bytes(ptr, 0, len) };
        Some(Self { ptr, len })
    }
    pub fn len(&self) -> usize { self.len }
    pub fn is_empty(&self) -> bool { self.len == 0 }
}
impl std::ops::Deref for PinnedBuffer {
    type Target = [f32];
    fn deref(&self) -> &[f32] { unsafe { std::slice::from_raw_parts(self.ptr, self.len) } }
}
impl std::ops::DerefMut for PinnedBuffer {
    fn deref_mut(&mut self) -> &mut [f32] { unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) } }
}
impl Drop for PinnedBuffer {
    fn drop(&mut self) { unsafe { rsb_free_pinned(self.ptr as *mut c_void) } }
}

/// `len` f32 values of device memory on one GPU (RAII over `rsb_alloc_device`), for callers that keep
/// their audio resident on the device between submits.
pub struct DeviceBuffer {
    ptr: *mut f32,
    len: usize,
    device: i32,
}
unsafe impl Send for DeviceBuffer {}
impl DeviceBuffer {
    pub fn new(device: i32, len: usize) -> Option<Self> {
        let ptr = unsafe { rsb_alloc_device(device, len * std::mem::size_of::<f32>()) } as *mut f32;
        if ptr.is_null() { None } else { Some(Self { ptr, len, device }) }
    }
    pub fn len(&self) -> usize { self.len }
    pub fn is_empty(&self) -> bool { self.len == 0 }
    /// synchronous host -> device copy into `[offset, offset + src.len())`
    pub fn upload(&mut self, offset: usize, src: &[f32]) {
        assert!(offset + src.len() <= self.len);
        let rc = unsafe {
            rsb_memcpy(self.device, self.ptr.add(offset) as *mut c_void, src.as_ptr() as *const c_void,
                       std::mem::size_of_val(src), 0)
        };
        assert!(rc == 0, "resampler-cuda: upload failed ({rc})");
    }
    /// synchronous device -> host copy of `[offset, offset + dst.len())`
    pub fn download(&self, offset: usize, dst: &mut [f32]) {
        assert!(offset + dst.len() <= self.len);
        let rc = unsafe {
            rsb_memcpy(self.device, dst.as_mut_ptr() as *mut c_void, self.ptr.add(offset) as *const c_void,
                       std::mem::size_of_val(dst), 1)
        };
        assert!(rc == 0, "resampler-cuda: download failed ({rc})");
    }
}
impl Drop for DeviceBuffer {
    fn drop(&mut self) { unsafe { rsb_free_device(self.device, self.ptr as *mut c_void) } }
}

/// Totals of one `process_batch` job.
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct BatchCounts {
    /// input values consumed (always the whole input unless the output capacity ran out)
    pub consumed: usize,
    /// output values written
    pub produced: usize,
    /// `resample()` calls the canonical loop made
    pub calls: u32,
}

/// Work submitted with [`FirBatch::submit_device_async`]; borrows the handle and the buffers until
/// [`PendingSubmit::sync`] (or drop) has waited for the GPU, so neither can be touched early.
pub struct PendingSubmit<'a> {
    batch: &'a mut FirBatch,
    counts: Vec<(usize, usize)>,
    consumed: Vec<usize>,
    produced: Vec<usize>,
    done: bool,
    _buffers: std::marker::PhantomData<&'a mut DeviceBuffer>,
}
impl PendingSubmit<'_> {
    /// Waits for the submit (and everything queued before it on the handle's CUDA stream) and
    /// returns `(consumed, produced)` per listed stream.
    pub fn sync(mut self) -> Result<Vec<(usize, usize)>, ResampleError> {
        self.done = true;
        map_err(unsafe { rsb_fir_sync(self.batch.h) })?;
        self.counts = self.consumed.iter().copied().zip(self.produced.iter().copied()).collect();
        Ok(std::mem::take(&mut self.counts))
    }
}
impl Drop for PendingSubmit<'_> {
    fn drop(&mut self) {
        if !self.done { unsafe { rsb_fir_sync(self.batch.h) }; }
    }
}

/// N independent streams with identical parameters on one GPU.
pub struct FirBatch {
    h: *mut RsbFir,
    channels: usize,
}
// Send audit: the handle owns its CUDA streams, events and device buffers and binds the device
// itself at the top of every entry point (`cudaSetDevice`), so it may move between threads; every
// mutating entry point takes `&mut self`, i.e. one caller at a time, which is all the library asks
// for (it is NOT `Sync`: the C side keeps per-handle host mirrors without locks).
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

    /// The canonical caller loop (`resample()` on `call_len` values at a time until the input is
    /// used up, resample/src/main.rs:226-254) for one job per stream, host slices in and out.
    /// Long equally sized jobs on pinned buffers ([`PinnedBuffer`]) run as a pipeline of time
    /// slices inside the library: host->device copy, kernels and device->host copy overlap.
    pub fn process_batch(&mut self, inputs: &[&[f32]], call_len: usize, outputs: &mut [&mut [f32]])
                         -> Result<Vec<BatchCounts>, ResampleError> {
        assert_eq!(inputs.len(), outputs.len());
        let n = inputs.len();
        let in_ptrs: Vec<*const f32> = inputs.iter().map(|s| s.as_ptr()).collect();
        let in_lens: Vec<usize> = inputs.iter().map(|s| s.len()).collect();
        let out_ptrs: Vec<*mut f32> = outputs.iter_mut().map(|s| s.as_mut_ptr()).collect();
        let out_caps: Vec<usize> = outputs.iter().map(|s| s.len()).collect();
        let (mut c, mut p, mut k) = (vec![0usize; n], vec![0usize; n], vec![0u32; n]);
        map_err(unsafe {
            rsb_fir_process_batch(self.h, n as u32, std::ptr::null(), in_ptrs.as_ptr(), in_lens.as_ptr(),
                                  call_len, 0, out_ptrs.as_ptr(), out_caps.as_ptr(), c.as_mut_ptr(),
                                  p.as_mut_ptr(), k.as_mut_ptr(), RSB_MEM_HOST, 0)
        })?;
        Ok((0..n).map(|i| BatchCounts { consumed: c[i], produced: p[i], calls: k[i] }).collect())
    }

    /// One `resample()` call per listed stream on DEVICE-resident buffers, enqueued on the handle's
    /// CUDA stream without waiting: `inputs[i]` / `outputs[i]` are `(buffer, offset, len)` in values.
    /// The returned guard keeps the handle and the buffers borrowed until it has been synchronised.
    pub fn submit_device_async<'a>(&'a mut self, streams: &[u32], inputs: &[(&'a DeviceBuffer, usize, usize)],
                                   outputs: &mut [(&'a mut DeviceBuffer, usize, usize)])
                                   -> Result<PendingSubmit<'a>, ResampleError> {
        assert!(streams.len() == inputs.len() && inputs.len() == outputs.len());
        let n = inputs.len();
        for (b, off, len) in inputs.iter() { assert!(off + len <= b.len()); }
        for (b, off, len) in outputs.iter() { assert!(off + len <= b.len()); }
        let in_ptrs: Vec<*const f32> = inputs.iter().map(|(b, off, _)| unsafe { b.ptr.add(*off) as *const f32 }).collect();
        let in_lens: Vec<usize> = inputs.iter().map(|(_, _, len)| *len).collect();
        let out_ptrs: Vec<*mut f32> = outputs.iter().map(|(b, off, _)| unsafe { b.ptr.add(*off) }).collect();
        let out_lens: Vec<usize> = outputs.iter().map(|(_, _, len)| *len).collect();
        let mut pending = PendingSubmit { batch: self, counts: Vec::new(), consumed: vec![0usize; n],
                                          produced: vec![0usize; n], done: false,
                                          _buffers: std::marker::PhantomData };
        // the count vectors live in the guard: the library writes them at sync() at the latest
        map_err(unsafe {
            rsb_fir_submit_batch(pending.batch.h, n as u32, streams.as_ptr(), in_ptrs.as_ptr(), in_lens.as_ptr(),
                                 out_ptrs.as_ptr(), out_lens.as_ptr(), pending.consumed.as_mut_ptr(),
                                 pending.produced.as_mut_ptr(), RSB_MEM_DEVICE, RSB_FLAG_ASYNC)
        })?;
        Ok(pending)
    }

    /// Waits for everything enqueued on this handle.
    pub fn sync(&mut self) -> Result<(), ResampleError> { map_err(unsafe { rsb_fir_sync(self.h) }) }

    /// The handle's `cudaStream_t` as an opaque pointer, for callers that order their own CUDA
    /// work (copies, kernels) against the submits.
    pub fn cuda_stream(&self) -> *mut c_void { unsafe { rsb_fir_cuda_stream(self.h) } }

    /// `(batches, slices)` of host-memspace `process_batch` calls that ran pipelined.
    pub fn host_pipeline_stats(&self) -> (u64, u64) {
        let (mut b, mut s) = (0u64, 0u64);
        unsafe { rsb_fir_host_pipeline_stats(self.h, &mut b, &mut s) };
        (b, s)
    }
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

/// Drop-in for `resampler::ResamplerFft` (src/resampler_fft.rs:43-246): fixed-size chunks.
pub struct ResamplerFft { h: *mut RsbFft }
unsafe impl Send for ResamplerFft {}

impl ResamplerFft {
    /// `ResamplerFft::new` (resampler_fft.rs:75-128)
    pub fn new(channels: usize, sample_rate_input: SampleRate, sample_rate_output: SampleRate) -> Self {
        let mut h = std::ptr::null_mut();
        let rc = unsafe {
            rsb_fft_create(&mut h, 0, 1, channels as u32, u32::from(sample_rate_input), u32::from(sample_rate_output))
        };
        assert!(rc == 0, "resampler-cuda: create failed ({rc})");
        Self { h }
    }
    /// resampler_fft.rs:135-137
    pub fn chunk_size_input(&self) -> usize { unsafe { rsb_fft_chunk_size_input(self.h) } }
    /// resampler_fft.rs:143-145
    pub fn chunk_size_output(&self) -> usize { unsafe { rsb_fft_chunk_size_output(self.h) } }
    /// resampler_fft.rs:151-153
    pub fn delay(&self) -> usize { unsafe { rsb_fft_delay(self.h) } }
    /// resampler_fft.rs:182-246
    pub fn resample(&mut self, input: &[f32], output: &mut [f32]) -> Result<(), ResampleError> {
        map_err(unsafe {
            rsb_fft_resample(self.h, 0, input.as_ptr(), input.len(), output.as_mut_ptr(), output.len())
        })
    }
}
impl Drop for ResamplerFft {
    fn drop(&mut self) { unsafe { rsb_fft_destroy(self.h) } }
}

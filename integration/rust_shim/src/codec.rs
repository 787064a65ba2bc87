//! Drop-in replacement for the reference's `src/codec.rs` (ajcm474/gapless-lossy-codec v0.5.0):
//! the same public items with the same signatures, implemented over the C ABI of libglc_b200.so
//! (`include/glc.h`).  `src/lib.rs`, `src/main.rs`, `src/audio.rs`, `src/ui.rs`, `src/playback.rs`
//! and `tests/*.rs` compile unchanged: both `lib.rs` and `main.rs` declare `mod codec; mod flac;`,
//! and the raw bindings live in the private submodule `codec::sys`, so no other file is touched.
//!
//! There is no CPU fallback: without a usable B200 every entry point fails (constructors panic with
//! the library's message, fallible functions return it as an `anyhow::Error`).
//!
//! The data model (EncodedAudio .. AudioChunk) is the crate's serde/bincode schema and is kept
//! field for field (reference src/codec.rs:31-85); everything below it is new.
use anyhow::{anyhow, Result};
use crossbeam_channel::{bounded, Receiver, Sender};
use serde::{Deserialize, Serialize};
use std::ffi::CStr;
use std::ptr;
use std::slice;
use std::sync::{Arc, Mutex, MutexGuard, OnceLock};
use std::time::Instant;

const HOP_SIZE: usize = 1024; // reference src/codec.rs:16
const FRAMES_PER_CHUNK: usize = 500; // reference src/codec.rs:18

// ------------------------------------------------------------------ data model (unchanged schema)

#[derive(Serialize, Deserialize, Debug, Clone)]
pub struct EncodedAudio
{
    pub header: AudioHeader,
    pub frames: Vec<EncodedFrame>,
    pub gapless_info: GaplessInfo,
}

#[derive(Serialize, Deserialize, Debug, Clone)]
pub struct AudioHeader
{
    pub sample_rate: u32,
    pub channels: u16,
    pub total_samples: u64,
}

#[derive(Serialize, Deserialize, Debug, Clone)]
pub struct GaplessInfo
{
    pub encoder_delay: u32,
    pub padding: u32,
    pub original_length: u64,
}

#[derive(Serialize, Deserialize, Debug, Clone)]
pub struct EncodedFrame
{
    pub sparse_coeffs_per_channel: Vec<Vec<(u16, i16)>>,
    pub scale_factors: Vec<f32>,
    pub raw_pcm: Option<Vec<i16>>,
}

pub enum Progress
{
    Encoding(f32),
    Decoding(f32),
    Exporting(f32),
    Complete(String),
    Error(String),
    Status(String),
}

pub struct AudioChunk
{
    pub samples: Vec<f32>,
    pub is_last: bool,
}

// ------------------------------------------------------------------ raw bindings (include/glc.h)

pub(crate) mod sys
{
    use std::os::raw::{c_char, c_int, c_void};

    #[repr(C)]
    #[derive(Clone, Copy)]
    pub struct GlcPair
    {
        pub idx: u16,
        pub q: i16,
    }

    /// struct glc_encoded: the flat image of EncodedAudio
    #[repr(C)]
    pub struct GlcEncoded
    {
        pub sample_rate: u32,
        pub channels: u16,
        pub reserved0: u16,
        pub total_samples: u64,
        pub encoder_delay: u32,
        pub padding: u32,
        pub original_length: u64,
        pub n_frames: u64,
        pub frame_is_raw: *mut u8,
        pub nnz: *mut u32,
        pub pair_offset: *mut u64,
        pub pairs: *mut GlcPair,
        pub scales: *mut f32,
        pub raw_offset: *mut u64,
        pub raw: *mut i16,
    }

    #[repr(C)]
    pub struct GlcCtx
    {
        _private: [u8; 0],
    }
    #[repr(C)]
    pub struct GlcEncoder
    {
        _private: [u8; 0],
    }
    #[repr(C)]
    pub struct GlcDecoder
    {
        _private: [u8; 0],
    }
    #[repr(C)]
    pub struct GlcStream
    {
        _private: [u8; 0],
    }

    pub const GLC_OK: c_int = 0;
    pub const GLC_ERR_TOO_SHORT: c_int = 2;
    pub const GLC_MODE_EXACT: c_int = 0;

    unsafe extern "C" {
        pub fn glc_last_error() -> *const c_char;
        pub fn glc_ctx_create(device: c_int, mode: c_int, out: *mut *mut GlcCtx) -> c_int;
        pub fn glc_free(ctx: *mut GlcCtx, p: *mut c_void);
        // Encoder::new / Encoder::encode                                  reference src/codec.rs:406, 421
        pub fn glc_encoder_new(ctx: *mut GlcCtx, sample_rate: u32, out: *mut *mut GlcEncoder) -> c_int;
        pub fn glc_encoder_free(enc: *mut GlcEncoder);
        pub fn glc_encode(enc: *mut GlcEncoder, pcm: *const f32, n_samples: u64, channels: u16,
                          out: *mut *mut GlcEncoded) -> c_int;
        pub fn glc_encoded_free(ctx: *mut GlcCtx, e: *mut GlcEncoded);
        // Decoder::new / decode / decode_streaming                          reference src/codec.rs:581, 744, 595
        pub fn glc_decoder_new(ctx: *mut GlcCtx, channels: u32, sample_rate: u32, out: *mut *mut GlcDecoder) -> c_int;
        pub fn glc_decoder_free(dec: *mut GlcDecoder);
        pub fn glc_decode(dec: *mut GlcDecoder, enc: *const GlcEncoded, pcm: *mut *mut f32, n_samples: *mut u64) -> c_int;
        pub fn glc_decode_stream_open(dec: *mut GlcDecoder, enc: *const GlcEncoded, out: *mut *mut GlcStream) -> c_int;
        pub fn glc_decode_stream_next(s: *mut GlcStream, samples: *mut *const f32, n_samples: *mut u64,
                                      is_last: *mut c_int, progress_percent: *mut f32) -> c_int;
        pub fn glc_decode_stream_close(s: *mut GlcStream);
        // flac::encode_flac_with_level                                      reference src/flac.rs:947
        pub fn glc_flac_encode(ctx: *mut GlcCtx, pcm: *const f32, n_samples: u64, sample_rate: u32, channels: u16,
                               level: u8, bytes: *mut *mut u8, len: *mut u64) -> c_int;
    }
}

use sys::*;

// ------------------------------------------------------------------ one context per process

/// A glc_ctx may be used by one thread at a time (include/glc.h); `cargo test` runs tests on many
/// threads and decode_streaming has its own, so every call into the library holds this lock.
pub(crate) struct Ctx(pub(crate) *mut GlcCtx);
unsafe impl Send for Ctx {}

static CTX: OnceLock<Mutex<Ctx>> = OnceLock::new();

pub(crate) fn last_error() -> String
{
    let p = unsafe { glc_last_error() };
    if p.is_null()
    {
        return String::from("libglc_b200: unknown error");
    }
    unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned()
}

pub(crate) fn check(status: i32) -> Result<()>
{
    if status == GLC_OK { Ok(()) } else { Err(anyhow!("{}", last_error())) }
}

/// Locks the process-wide context, creating it on first use (device from GLC_DEVICE, default 0).
/// Panics when no B200 is usable: there is no CPU fallback.
pub(crate) fn ctx() -> MutexGuard<'static, Ctx>
{
    let m = CTX.get_or_init(||
    {
        let device: i32 = std::env::var("GLC_DEVICE").ok().and_then(|v| v.parse().ok()).unwrap_or(0);
        let mut h: *mut GlcCtx = ptr::null_mut();
        let st = unsafe { glc_ctx_create(device, GLC_MODE_EXACT, &mut h) };
        if st != GLC_OK
        {
            panic!("libglc_b200: {}", last_error());
        }
        Mutex::new(Ctx(h))
    });
    m.lock().unwrap_or_else(|poisoned| poisoned.into_inner())
}

// ------------------------------------------------------------------ nested <-> flat

/// EncodedAudio (nested Vecs) -> the flat arrays struct glc_encoded points into.
#[allow(dead_code)] // the Vecs are only kept alive for `view`
struct Flat
{
    frame_is_raw: Vec<u8>,
    nnz: Vec<u32>,
    pair_offset: Vec<u64>,
    pairs: Vec<GlcPair>,
    scales: Vec<f32>,
    raw_offset: Vec<u64>,
    raw: Vec<i16>,
    view: GlcEncoded,
}
unsafe impl Send for Flat {}

impl Flat
{
    fn new(e: &EncodedAudio) -> Box<Flat>
    {
        let frames = e.frames.len();
        let ch = e.header.channels as usize;
        let n_pairs: usize = e.frames.iter().map(|f| f.sparse_coeffs_per_channel.iter().map(|v| v.len()).sum::<usize>()).sum();
        let n_raw: usize = e.frames.iter().map(|f| f.raw_pcm.as_ref().map_or(0, |r| r.len())).sum();
        let mut frame_is_raw: Vec<u8> = Vec::with_capacity(frames);
        let mut nnz: Vec<u32> = Vec::with_capacity(frames * ch);
        let mut pair_offset: Vec<u64> = Vec::with_capacity(frames * ch + 1);
        let mut pairs: Vec<GlcPair> = Vec::with_capacity(n_pairs);
        let mut scales: Vec<f32> = Vec::with_capacity(frames * ch);
        let mut raw_offset: Vec<u64> = Vec::with_capacity(frames + 1);
        let mut raw: Vec<i16> = Vec::with_capacity(n_raw);
        for f in &e.frames
        {
            raw_offset.push(raw.len() as u64);
            match &f.raw_pcm
            {
                Some(r) =>
                {
                    // reference src/codec.rs:626-644: a frame with raw_pcm ignores its sparse vectors
                    frame_is_raw.push(1);
                    raw.extend_from_slice(r);
                    for _ in 0..ch
                    {
                        pair_offset.push(pairs.len() as u64);
                        nnz.push(0);
                        scales.push(0.0);
                    }
                }
                None =>
                {
                    frame_is_raw.push(0);
                    for c in 0..ch
                    {
                        pair_offset.push(pairs.len() as u64);
                        let empty: &[(u16, i16)] = &[];
                        let v: &[(u16, i16)] = f.sparse_coeffs_per_channel.get(c).map(|v| v.as_slice()).unwrap_or(empty);
                        nnz.push(v.len() as u32);
                        pairs.extend(v.iter().map(|&(idx, q)| GlcPair { idx, q }));
                        scales.push(f.scale_factors.get(c).copied().unwrap_or(0.0));
                    }
                }
            }
        }
        pair_offset.push(pairs.len() as u64);
        raw_offset.push(raw.len() as u64);
        // the Vecs' heap buffers do not move when the Vecs are moved into the Box
        let view = GlcEncoded
        {
            sample_rate: e.header.sample_rate,
            channels: e.header.channels,
            reserved0: 0,
            total_samples: e.header.total_samples,
            encoder_delay: e.gapless_info.encoder_delay,
            padding: e.gapless_info.padding,
            original_length: e.gapless_info.original_length,
            n_frames: frames as u64,
            frame_is_raw: frame_is_raw.as_mut_ptr(),
            nnz: nnz.as_mut_ptr(),
            pair_offset: pair_offset.as_mut_ptr(),
            pairs: pairs.as_mut_ptr(),
            scales: scales.as_mut_ptr(),
            raw_offset: raw_offset.as_mut_ptr(),
            raw: raw.as_mut_ptr(),
        };
        Box::new(Flat { frame_is_raw, nnz, pair_offset, pairs, scales, raw_offset, raw, view })
    }
}

/// `slice::from_raw_parts` must not see a null pointer, not even for an empty slice.
unsafe fn view<'a, T>(p: *const T, n: usize) -> &'a [T]
{
    if n == 0 || p.is_null() { &[] } else { unsafe { slice::from_raw_parts(p, n) } }
}

/// struct glc_encoded -> EncodedAudio, frame by frame as Encoder::encode builds them
/// (reference src/codec.rs:521-564).
unsafe fn nested_from_flat(e: &GlcEncoded) -> EncodedAudio
{
    let frames = e.n_frames as usize;
    let ch = e.channels as usize;
    let (raw_flag, po, ro, scales): (&[u8], &[u64], &[u64], &[f32]) = unsafe
    {
        (view(e.frame_is_raw, frames), view(e.pair_offset, frames * ch + 1), view(e.raw_offset, frames + 1),
         view(e.scales, frames * ch))
    };
    let n_pairs = po[frames * ch] as usize;
    let n_raw = ro[frames] as usize;
    let pairs: &[GlcPair] = unsafe { view(e.pairs, n_pairs) };
    let raw: &[i16] = unsafe { view(e.raw, n_raw) };
    let mut out = Vec::with_capacity(frames);
    for f in 0..frames
    {
        if raw_flag[f] != 0
        {
            out.push(EncodedFrame
            {
                sparse_coeffs_per_channel: vec![],
                scale_factors: vec![],
                raw_pcm: Some(raw[ro[f] as usize..ro[f + 1] as usize].to_vec()),
            });
        }
        else
        {
            let sparse = (0..ch).map(|c|
            {
                let (a, b) = (po[f * ch + c] as usize, po[f * ch + c + 1] as usize);
                pairs[a..b].iter().map(|p| (p.idx, p.q)).collect::<Vec<(u16, i16)>>()
            }).collect();
            out.push(EncodedFrame
            {
                sparse_coeffs_per_channel: sparse,
                scale_factors: scales[f * ch..(f + 1) * ch].to_vec(),
                raw_pcm: None,
            });
        }
    }
    EncodedAudio
    {
        header: AudioHeader { sample_rate: e.sample_rate, channels: e.channels, total_samples: e.total_samples },
        frames: out,
        gapless_info: GaplessInfo { encoder_delay: e.encoder_delay, padding: e.padding, original_length: e.original_length },
    }
}

// ------------------------------------------------------------------ Encoder

pub struct Encoder
{
    h: *mut GlcEncoder,
}
unsafe impl Send for Encoder {}

impl Encoder
{
    /// reference src/codec.rs:406
    pub fn new(sample_rate: u32) -> Self
    {
        let c = ctx();
        let mut h: *mut GlcEncoder = ptr::null_mut();
        let st = unsafe { glc_encoder_new(c.0, sample_rate, &mut h) };
        if st != GLC_OK
        {
            panic!("libglc_b200: {}", last_error());
        }
        Self { h }
    }

    /// reference src/codec.rs:421
    pub fn encode(&mut self, samples: &[f32], channels: u16) -> Result<EncodedAudio>
    {
        let c = ctx();
        let mut e: *mut GlcEncoded = ptr::null_mut();
        let st = unsafe { glc_encode(self.h, samples.as_ptr(), samples.len() as u64, channels, &mut e) };
        if st == GLC_ERR_TOO_SHORT
        {
            // <= 512 samples per channel: the reference panics (src/codec.rs:449-452 then :474)
            panic!("range end index out of range for slice ({})", last_error());
        }
        check(st)?;
        let out = unsafe { nested_from_flat(&*e) };
        unsafe { glc_encoded_free(c.0, e) };
        Ok(out)
    }
}

impl Drop for Encoder
{
    fn drop(&mut self)
    {
        let _c = ctx();
        unsafe { glc_encoder_free(self.h) };
    }
}

// ------------------------------------------------------------------ Decoder

pub struct Decoder
{
    h: Arc<DecoderHandle>,
}

struct DecoderHandle(*mut GlcDecoder);
unsafe impl Send for DecoderHandle {}
unsafe impl Sync for DecoderHandle {}
impl Drop for DecoderHandle
{
    fn drop(&mut self)
    {
        let _c = ctx();
        unsafe { glc_decoder_free(self.0) };
    }
}

impl Decoder
{
    /// reference src/codec.rs:581 (both arguments are informational there too: the header wins, :598)
    pub fn new(channels: usize, sample_rate: u32) -> Self
    {
        let c = ctx();
        let mut h: *mut GlcDecoder = ptr::null_mut();
        let st = unsafe { glc_decoder_new(c.0, channels as u32, sample_rate, &mut h) };
        if st != GLC_OK
        {
            panic!("libglc_b200: {}", last_error());
        }
        Self { h: Arc::new(DecoderHandle(h)) }
    }

    /// reference src/codec.rs:595: one thread, a bounded(5) channel, chunks of exactly 500 frames and
    /// a last chunk with is_last; Progress::Status / Decoding(%) / Complete as the reference sends them.
    pub fn decode_streaming(&mut self, encoded: Arc<EncodedAudio>, progress_sender: Option<Sender<Progress>>) -> Receiver<AudioChunk>
    {
        let (tx, rx) = bounded(5);
        let handle = self.h.clone();
        std::thread::spawn(move ||
        {
            let start_time = Instant::now();
            let total_frames = encoded.frames.len();
            if let Some(ref s) = progress_sender
            {
                let _ = s.send(Progress::Status(format!("Starting streaming decode of {} frames", total_frames)));
            }
            let flat = Flat::new(&encoded);
            let mut stream: *mut GlcStream = ptr::null_mut();
            {
                let _c = ctx();
                if unsafe { glc_decode_stream_open(handle.0, &flat.view, &mut stream) } != GLC_OK
                {
                    if let Some(ref s) = progress_sender
                    {
                        let _ = s.send(Progress::Error(last_error()));
                    }
                    return;
                }
            }
            loop
            {
                let mut p: *const f32 = ptr::null();
                let mut n: u64 = 0;
                let mut last: i32 = 0;
                let mut pct: f32 = 0.0;
                let chunk =
                {
                    let _c = ctx();
                    if unsafe { glc_decode_stream_next(stream, &mut p, &mut n, &mut last, &mut pct) } != GLC_OK
                    {
                        if let Some(ref s) = progress_sender
                        {
                            let _ = s.send(Progress::Error(last_error()));
                        }
                        break;
                    }
                    // the chunk memory belongs to the stream until the next call: copy it out under the lock
                    AudioChunk { samples: unsafe { view(p, n as usize) }.to_vec(), is_last: last != 0 }
                };
                if last == 0
                {
                    if let Some(ref s) = progress_sender
                    {
                        let _ = s.send(Progress::Decoding(pct));
                    }
                }
                let _ = tx.send(chunk);
                if last != 0
                {
                    break;
                }
            }
            {
                let _c = ctx();
                unsafe { glc_decode_stream_close(stream) };
            }
            if let Some(ref s) = progress_sender
            {
                let _ = s.send(Progress::Complete(format!("Decoded {} frames in {:.2}s", total_frames, start_time.elapsed().as_secs_f32())));
            }
        });
        rx
    }

    /// reference src/codec.rs:744: all chunks, then the gapless trim (drop encoder_delay interleaved
    /// VALUES, truncate to original_length) -- one device pass here, the trim applied inside glc_decode.
    pub fn decode(&mut self, encoded: &EncodedAudio, progress_sender: Option<Sender<Progress>>) -> Result<Vec<f32>>
    {
        let start_time = Instant::now();
        let total_frames = encoded.frames.len();
        if let Some(ref s) = progress_sender
        {
            let _ = s.send(Progress::Status(format!("Starting streaming decode of {} frames", total_frames)));
        }
        let flat = Flat::new(encoded);
        let c = ctx();
        let mut p: *mut f32 = ptr::null_mut();
        let mut n: u64 = 0;
        check(unsafe { glc_decode(self.h.0, &flat.view, &mut p, &mut n) })?;
        let all = unsafe { view(p as *const f32, n as usize) }.to_vec();
        unsafe { glc_free(c.0, p as *mut std::os::raw::c_void) };
        drop(c);
        if let Some(ref s) = progress_sender
        {
            // the values the reference sends before each full chunk (src/codec.rs:708-714)
            let mut idx = FRAMES_PER_CHUNK;
            while idx <= total_frames
            {
                let _ = s.send(Progress::Decoding(((idx - 1) as f32) / (total_frames as f32) * 100.0));
                idx += FRAMES_PER_CHUNK;
            }
            let _ = s.send(Progress::Complete(format!("Decoded {} frames in {:.2}s", total_frames, start_time.elapsed().as_secs_f32())));
        }
        debug_assert!(all.len() <= (total_frames + 1) * HOP_SIZE * encoded.header.channels as usize);
        Ok(all)
    }
}

// ------------------------------------------------------------------ container (unchanged: bincode)

/// reference src/codec.rs:774
pub fn save_encoded(encoded: &EncodedAudio, path: &std::path::Path) -> Result<()>
{
    let data = bincode::serialize(encoded)?;
    std::fs::write(path, data)?;
    Ok(())
}

/// reference src/codec.rs:781
pub fn load_encoded(path: &std::path::Path) -> Result<EncodedAudio>
{
    let data = std::fs::read(path)?;
    let encoded = bincode::deserialize(&data)?;
    Ok(encoded)
}

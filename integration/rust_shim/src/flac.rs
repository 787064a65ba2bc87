//! Drop-in replacement for the reference's `src/flac.rs` (public items of reference
//! src/flac.rs:947-1088) over libglc_b200.so.  Bytes are identical to the reference encoder's: fixed
//! predictors, per-partition Rice parameter, 16-bit samples, independent channels.
use anyhow::Result;
use std::path::Path;
use std::ptr;
use std::slice;

use crate::codec::sys::{glc_flac_encode, glc_free};
use crate::codec::{check, ctx};

/// reference src/flac.rs:947.  Errors keep the reference's order and wording: "< 16 samples per
/// channel" first (:963-969), then "Invalid compression level" (:972-978).
pub fn encode_flac_with_level(samples: &[f32], sample_rate: u32, channels: u16, compression_level: u8) -> Result<Vec<u8>>
{
    let c = ctx();
    let mut p: *mut u8 = ptr::null_mut();
    let mut n: u64 = 0;
    check(unsafe
    {
        glc_flac_encode(c.0, samples.as_ptr(), samples.len() as u64, sample_rate, channels, compression_level, &mut p, &mut n)
    })?;
    let bytes = if n == 0 || p.is_null() { Vec::new() } else { unsafe { slice::from_raw_parts(p, n as usize) }.to_vec() };
    unsafe { glc_free(c.0, p as *mut std::os::raw::c_void) };
    Ok(bytes)
}

/// reference src/flac.rs:1055 (level 5)
pub fn encode_flac(samples: &[f32], sample_rate: u32, channels: u16) -> Result<Vec<u8>>
{
    encode_flac_with_level(samples, sample_rate, channels, 5)
}

/// reference src/flac.rs:1065
pub fn export_to_flac_with_level(path: &Path, samples: &[f32], sample_rate: u32, channels: u16, compression_level: u8) -> Result<()>
{
    let flac_data = encode_flac_with_level(samples, sample_rate, channels, compression_level)?;
    std::fs::write(path, flac_data)?;
    Ok(())
}

/// reference src/flac.rs:1080
pub fn export_to_flac(path: &Path, samples: &[f32], sample_rate: u32, channels: u16) -> Result<()>
{
    export_to_flac_with_level(path, samples, sample_rate, channels, 5)
}

// dump_reference.rs -- PINNING KIT for the oracle (and through it, the CUDA path).
//
// Not built here: this image has no cargo/rustc.  On any machine with a Rust toolchain:
//
//     cp tests/golden/dump_reference.rs  <reference checkout>/tests/dump_reference.rs
//     cd <reference checkout>
//     GLC_DUMP_DIR=/tmp/ref_v1 cargo test --release --test dump_reference -- --nocapture
//     cp -r /tmp/ref_v1  <this repo>/tests/golden/ref_v1
//     python -m pytest tests/test_reference_golden.py            # oracle vs the real crate
//     python -m pytest tests/test_reference_golden.py -m gpu     # CUDA path vs the real crate (B200)
//
// It is a Cargo integration test that uses only the crate's PUBLIC API (src/lib.rs:1-5):
// Encoder::encode, Decoder::decode, Decoder::decode_streaming, codec::save_encoded,
// flac::encode_flac_with_level.  For every case it writes, into $GLC_DUMP_DIR:
//
//     <name>.in.f32      the input samples it generated (interleaved f32, little endian) -- the
//                        comparer feeds exactly these bytes to the oracle / the CUDA path, so no
//                        cross-language generator has to agree
//     <name>.glc         save_encoded(Encoder::new(rate).encode(samples, ch))   (bincode image)
//     <name>.pcm.f32     Decoder::decode of that stream (gapless-trimmed)
//     <name>.stream.f32  the chunks of Decoder::decode_streaming concatenated (untrimmed)
//     <name>.chunks.txt  one line per chunk: "<len> <is_last>"
//     <name>.l<N>.flac   encode_flac_with_level(samples, rate, ch, N)
//     manifest.txt       one line per artefact set: kind name sample_rate channels [levels]
//     table.fnv          FNV-1a/64 of the f32 bit patterns of the cosine table and window as
//                        MdctTables::new computes them (src/codec.rs:326-356 restated with public f32
//                        ops): tells the comparer whether dump host and test host share a libm
//
// Inputs come from the reference's own generators (tests/utils.rs).
use gapless_lossy_codec::codec::{save_encoded, Decoder, Encoder};
use gapless_lossy_codec::flac::encode_flac_with_level;
use std::f32::consts::PI;
use std::fs;
use std::io::Write;
use std::path::{Path, PathBuf};
use std::sync::Arc;

mod utils;
use utils::{generate_frequency_sweep, generate_sawtooth_wave, generate_sine_wave, generate_square_wave,
            generate_white_noise};

fn dump_dir() -> PathBuf
{
    let d = std::env::var("GLC_DUMP_DIR").unwrap_or_else(|_| "target/ref_v1".to_string());
    fs::create_dir_all(&d).expect("create dump dir");
    PathBuf::from(d)
}

fn write_f32(path: &Path, v: &[f32])
{
    let mut bytes = Vec::with_capacity(v.len() * 4);
    for x in v
    {
        bytes.extend_from_slice(&x.to_le_bytes());
    }
    fs::write(path, bytes).expect("write f32 file");
}

/// 70 % tones + low-passed noise, 30 % white noise per second: sparse AND raw frames in one file
/// (the shape of SURVEY.md 8(d) item 2, built from the reference's generators only).
fn music_like(sample_rate: u32, channels: u16, seconds: f32, seed: u64) -> Vec<f32>
{
    let ch = channels as usize;
    let tone_a = generate_sine_wave(440.0, sample_rate, channels, seconds);
    let tone_b = generate_sawtooth_wave(173.0, sample_rate, channels, seconds);
    let sweep = generate_frequency_sweep(100.0, 6000.0, sample_rate, channels, seconds);
    let noise = generate_white_noise(sample_rate, channels, seconds, seed);
    let total = tone_a.len() / ch;
    let mut out = vec![0.0f32; total * ch];
    for i in 0..total
    {
        let pos = (i % sample_rate as usize) as f32 / sample_rate as f32;
        for c in 0..ch
        {
            let k = i * ch + c;
            out[k] = if pos >= 0.7
            {
                noise[k]
            }
            else
            {
                0.4 * tone_a[k] + 0.3 * tone_b[k] + 0.3 * sweep[k] + 0.05 * noise[k]
            };
        }
    }
    out
}

fn fnv1a(bits: impl Iterator<Item = u32>) -> u64
{
    let mut h: u64 = 0xcbf29ce484222325;
    for w in bits
    {
        for b in w.to_le_bytes()
        {
            h ^= b as u64;
            h = h.wrapping_mul(0x100000001b3);
        }
    }
    h
}

/// MdctTables::new restated with the same public f32 operations (src/codec.rs:326-356).
fn table_fingerprint() -> (u64, u64)
{
    let n: usize = 1024;
    let two_n = 2 * n;
    let scale = PI / n as f32;
    let mut cos_bits = Vec::with_capacity(n * two_n);
    for k in 0..n
    {
        for i in 0..two_n
        {
            let angle = scale * (i as f32 + 0.5 + n as f32 / 2.0) * (k as f32 + 0.5);
            cos_bits.push(angle.cos().to_bits());
        }
    }
    let win = (0..two_n).map(|i| (PI * (i as f32 + 0.5) / two_n as f32).sin().to_bits());
    (fnv1a(cos_bits.into_iter()), fnv1a(win))
}

fn dump_codec(dir: &Path, manifest: &mut String, name: &str, samples: &[f32], sample_rate: u32, channels: u16)
{
    write_f32(&dir.join(format!("{name}.in.f32")), samples);
    let mut encoder = Encoder::new(sample_rate);
    let encoded = encoder.encode(samples, channels).expect("encode");
    save_encoded(&encoded, &dir.join(format!("{name}.glc"))).expect("save_encoded");
    let mut decoder = Decoder::new(channels as usize, sample_rate);
    let decoded = decoder.decode(&encoded, None).expect("decode");
    assert_eq!(decoded.len(), samples.len());
    write_f32(&dir.join(format!("{name}.pcm.f32")), &decoded);
    // the streaming view: chunk shapes + untrimmed samples
    let rx = decoder.decode_streaming(Arc::new(encoded.clone()), None);
    let mut all = Vec::new();
    let mut chunks = String::new();
    while let Ok(chunk) = rx.recv()
    {
        chunks.push_str(&format!("{} {}\n", chunk.samples.len(), chunk.is_last as u8));
        all.extend_from_slice(&chunk.samples);
        if chunk.is_last
        {
            break;
        }
    }
    write_f32(&dir.join(format!("{name}.stream.f32")), &all);
    fs::write(dir.join(format!("{name}.chunks.txt")), chunks).expect("write chunks");
    manifest.push_str(&format!("codec {name} {sample_rate} {channels}\n"));
}

fn dump_flac(dir: &Path, manifest: &mut String, name: &str, samples: &[f32], sample_rate: u32, channels: u16,
             levels: &[u8])
{
    write_f32(&dir.join(format!("{name}.in.f32")), samples);
    let mut line = format!("flac {name} {sample_rate} {channels}");
    for &level in levels
    {
        let bytes = encode_flac_with_level(samples, sample_rate, channels, level).expect("flac encode");
        fs::write(dir.join(format!("{name}.l{level}.flac")), bytes).expect("write flac");
        line.push_str(&format!(" {level}"));
    }
    line.push('\n');
    manifest.push_str(&line);
}

#[test]
fn dump_reference_outputs()
{
    let dir = dump_dir();
    let mut manifest = String::new();

    // ---- codec (src/codec.rs): the reference's own test shapes + mixed / multichannel / ragged ones
    dump_codec(&dir, &mut manifest, "sine440_mono_2s", &generate_sine_wave(440.0, 44100, 1, 2.0), 44100, 1);
    dump_codec(&dir, &mut manifest, "sine440_stereo_2s", &generate_sine_wave(440.0, 44100, 2, 2.0), 44100, 2);
    dump_codec(&dir, &mut manifest, "square_mono_1s", &generate_square_wave(440.0, 44100, 1, 1.0), 44100, 1);
    dump_codec(&dir, &mut manifest, "saw_mono_1s", &generate_sawtooth_wave(440.0, 44100, 1, 1.0), 44100, 1);
    dump_codec(&dir, &mut manifest, "sweep_stereo_48k", &generate_frequency_sweep(100.0, 8000.0, 48000, 2, 1.0), 48000, 2);
    dump_codec(&dir, &mut manifest, "noise_stereo_raw", &generate_white_noise(44100, 2, 0.5, 12345), 44100, 2);
    dump_codec(&dir, &mut manifest, "music_stereo_3s", &music_like(44100, 2, 3.0, 12345), 44100, 2);
    dump_codec(&dir, &mut manifest, "music_6ch_48k", &music_like(48000, 6, 1.5, 1000), 48000, 6);
    dump_codec(&dir, &mut manifest, "sine_96k_mono", &generate_sine_wave(1000.0, 96000, 1, 0.5), 96000, 1);
    dump_codec(&dir, &mut manifest, "sine_8k_mono", &generate_sine_wave(300.0, 8000, 1, 1.0), 8000, 1);
    dump_codec(&dir, &mut manifest, "long_stream_12s", &music_like(44100, 1, 12.0, 7), 44100, 1); // > 500 frames: 2 chunks
    let ragged = generate_sine_wave(440.0, 44100, 1, 0.2);
    for n in [513usize, 1024, 1025, 1536, 1537, 4097]
    {
        dump_codec(&dir, &mut manifest, &format!("ragged_{n}"), &ragged[..n], 44100, 1);
    }

    // ---- flac (src/flac.rs): every level, tails, multichannel, long streams (3-byte frame numbers)
    let all_levels: Vec<u8> = (0u8..=8).collect();
    dump_flac(&dir, &mut manifest, "flac_music_stereo", &music_like(44100, 2, 0.6, 5), 44100, 2, &all_levels);
    dump_flac(&dir, &mut manifest, "flac_sine_1000", &generate_sine_wave(440.0, 44100, 1, 0.1)[..1000], 44100, 1, &all_levels);
    dump_flac(&dir, &mut manifest, "flac_noise", &generate_white_noise(44100, 1, 0.2, 99), 44100, 1, &[0, 2, 5, 8]);
    dump_flac(&dir, &mut manifest, "flac_6ch_48k", &music_like(48000, 6, 0.3, 3), 48000, 6, &[5, 8]);
    dump_flac(&dir, &mut manifest, "flac_96k_stereo", &music_like(96000, 2, 2.0, 7), 96000, 2, &[8]);
    dump_flac(&dir, &mut manifest, "flac_min16", &generate_sine_wave(440.0, 8000, 1, 0.002)[..16], 8000, 1, &[0, 5, 8]);
    let tails = generate_white_noise(44100, 1, 0.3, 42);
    for n in [17usize, 19, 20, 1151, 1153, 4097, 4100]
    {
        dump_flac(&dir, &mut manifest, &format!("flac_tail_{n}"), &tails[..n], 44100, 1, &[2, 5, 8]);
    }
    // 2 100 blocks of 1 152 samples: frame numbers >= 0x800 take the 3-byte coded form (src/flac.rs:438-443)
    dump_flac(&dir, &mut manifest, "flac_long_l1", &generate_sine_wave(330.0, 44100, 1, 55.0), 44100, 1, &[1]);

    let (cos_fnv, win_fnv) = table_fingerprint();
    fs::write(dir.join("table.fnv"), format!("{cos_fnv:016x} {win_fnv:016x}\n")).expect("write table.fnv");
    let mut f = fs::File::create(dir.join("manifest.txt")).expect("manifest");
    f.write_all(manifest.as_bytes()).expect("manifest");
    println!("wrote {} manifest lines to {}", manifest.lines().count(), dir.display());
}

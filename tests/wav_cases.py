"""WAV files for the parity tests of the firmware's test-signal reader (Utility.cpp:773-888), built byte by byte so
that the format-chunk variants the reader distinguishes (16 / 18 / 40 bytes, PCM / mono / 16 bit) can be produced."""
import struct

import numpy as np


def wav_bytes(samples, fmt_size=16, audio_format=1, channels=1, bits=16, rate=8000):
    data = np.asarray(samples, np.int16).tobytes()
    block_align = channels * bits // 8
    fmt = struct.pack("<HHIIHH", audio_format, channels, rate, rate * block_align, block_align, bits)
    fmt += b"\x00" * (fmt_size - 16)                      # cbSize / extension (zeros)
    body = b"WAVE" + b"fmt " + struct.pack("<I", fmt_size) + fmt + b"data" + struct.pack("<I", len(data)) + data
    return b"RIFF" + struct.pack("<I", len(body)) + body


def cases():
    """name -> (file bytes or None, num_samples limit, chunk size)"""
    r = np.random.Generator(np.random.PCG64(0x7414))
    s5000 = r.integers(-32768, 32768, 5000)
    s777 = r.integers(-32768, 32768, 777)
    return {
        "pcm16_fmt16_chunk128": (wav_bytes(s5000), 100000, 128),
        "pcm16_fmt16_chunk86": (wav_bytes(s5000), 100000, 86),
        "pcm16_fmt18": (wav_bytes(s777, fmt_size=18), 1000, 128),
        "pcm16_fmt40": (wav_bytes(s777, fmt_size=40), 1000, 86),
        "extensible_tag": (wav_bytes(s777, fmt_size=40, audio_format=0xFFFE), 1000, 128),
        "stereo": (wav_bytes(s777, channels=2), 1000, 128),
        "eight_bit": (wav_bytes(s777, bits=8), 1000, 128),
        "fmt20": (wav_bytes(s777, fmt_size=20), 1000, 128),
        "too_long": (wav_bytes(s5000), 4999, 128),
        "exact_limit": (wav_bytes(s5000), 5000, 128),
        "missing": (None, 1000, 128),
    }


def run(load, read, path, data, limit, chunk):
    """Drive a reader: load(path, limit) -> rc; read(chunk) -> float32 array or None.  Returns (rc, n_reads, samples
    of the reads that lay wholly inside the file, valid prefix of the read that ran past its end)."""
    rc = load(path, limit)
    if rc != 0:
        return rc, 0, np.zeros(0, np.float32), np.zeros(0, np.float32)
    size = len(data)
    pos = None
    full, tail, n = [], np.zeros(0, np.float32), 0
    # the data chunk starts after the 12-byte RIFF header, the 8 + fmt_size format chunk and the 8-byte data header
    fmt_size = int.from_bytes(data[16:20], "little")
    pos = 12 + 8 + fmt_size + 8
    while True:
        got = read(chunk)
        if got is None:
            break
        n += 1
        if pos + 2 * chunk <= size:
            full.append(got.copy())
        else:
            tail = got[: max(0, (size - pos) // 2)].copy()
        pos += 2 * chunk
        assert n < 100000
    return rc, n, (np.concatenate(full) if full else np.zeros(0, np.float32)), tail

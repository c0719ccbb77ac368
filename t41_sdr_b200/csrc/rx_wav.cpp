/*
 * rx_wav.cpp — the firmware's WAV test-signal reader behind the C-ABI (include/t41rx.h), host only.
 *
 * Mirrors load_wav() / readWave() of the reference (software/T41_SDR/Utility.cpp:773-888): 16-bit mono PCM files,
 * format chunk of 16, 18 or 40 bytes, samples scaled by 1 / 32768.  The reference keeps one open file in globals
 * (Utility.cpp:751-753); here it is a handle.
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "../../include/t41rx.h"

struct t41rx_wav {
  FILE *f = nullptr;
  unsigned long position = 0, size_wav = 0;   /* Utility.cpp:752 */
  uint16_t bits_per_sample = 0;               /* Utility.cpp:753 */
  uint32_t sample_rate = 0;
  uint32_t data_bytes = 0;
};

namespace {
uint16_t ReadU16(FILE *f) {
  uint16_t v = 0;
  if (fread(&v, sizeof(v), 1, f) != 1) v = 0;
  return v;
}
uint32_t ReadU32(FILE *f) {
  uint32_t v = 0;
  if (fread(&v, sizeof(v), 1, f) != 1) v = 0;
  return v;
}
}  // namespace

extern "C" {

/* Utility.cpp:773-864.  Returns the reference's codes: 0 ok, -1 cannot open, -2 format chunk size not 16 / 18 / 40,
   -3 not PCM / mono / 16 bit, -4 more than num_samples samples in the data chunk. */
int t41rx_load_wav(t41rx_wav **out, const char *input_file, uint32_t num_samples) {
  if (out) *out = nullptr;
  if (!out || !input_file) return -1;
  FILE *f = fopen(input_file, "rb");
  if (!f) return -1;
  t41rx_wav w;
  w.f = f;
  fseek(f, 0, SEEK_END);
  w.size_wav = (unsigned long)ftell(f);
  fseek(f, 0, SEEK_SET);
  char tmp[16];
  /* master RIFF chunk: id, size, format; then the format chunk's id (none of the four is checked by the reference) */
  size_t got = fread(tmp, 1, 4, f);
  (void)ReadU32(f);
  got += fread(tmp, 1, 4, f);
  got += fread(tmp, 1, 4, f);
  (void)got;
  const uint32_t sub1 = ReadU32(f);
  if (!(sub1 == 16 || sub1 == 18 || sub1 == 40)) {
    fclose(f);
    return -2;
  }
  const uint16_t audio_format = ReadU16(f), num_channels = ReadU16(f);
  w.sample_rate = ReadU32(f);
  (void)ReadU32(f);                           /* byteRate */
  const uint16_t block_align = ReadU16(f);
  w.bits_per_sample = ReadU16(f);
  if (audio_format != 1 || num_channels != 1 || w.bits_per_sample != 16) {
    fclose(f);
    return -3;
  }
  if (sub1 == 18) fseek(f, 38, SEEK_SET);     /* skip the extension */
  else if (sub1 == 40) fseek(f, 60, SEEK_SET);
  got = fread(tmp, 1, 4, f);                  /* "data" (not checked) */
  w.data_bytes = ReadU32(f);
  if (block_align == 0 || w.data_bytes / block_align > num_samples) {
    fclose(f);
    return -4;
  }
  w.position = (unsigned long)ftell(f);
  t41rx_wav *h = new (std::nothrow) t41rx_wav(w);
  if (!h) {
    fclose(f);
    return -1;
  }
  *out = h;
  return 0;
}

/* Utility.cpp:866-888.  1 = buf holds size_buf samples / 32768; 0 = end of file reached, file closed.  The end
   test is the reference's own (byte position + SAMPLE count against the file size, "likely missing the end of
   file here"): the last reads before it may run past the data; the reference then converts whatever its stack
   buffer held, here the unread tail is zero. */
int t41rx_read_wave(t41rx_wav *w, float *buf, int size_buf) {
  if (!w || !w->f || !buf || size_buf <= 0) return 0;
  const unsigned long current = (unsigned long)ftell(w->f);
  if (current + (unsigned long)size_buf >= w->size_wav) {
    fclose(w->f);
    w->f = nullptr;
    return 0;
  }
  const size_t want = (size_t)size_buf * w->bits_per_sample / 8;
  int16_t raw[512];
  int done = 0;
  while (done < size_buf) {
    const int n = (size_buf - done) < 512 ? (size_buf - done) : 512;
    memset(raw, 0, sizeof(raw));
    const size_t r = fread(raw, 1, (size_t)n * 2, w->f);
    for (int i = 0; i < n; ++i) buf[done + i] = (float)raw[i] / 32768.0f;
    done += n;
    if (r < (size_t)n * 2) {                  /* ran past the end: the rest stays zero */
      for (int i = done; i < size_buf; ++i) buf[i] = 0.0f;
      break;
    }
  }
  (void)want;
  return 1;
}

uint32_t t41rx_wav_sample_rate(const t41rx_wav *w) { return w ? w->sample_rate : 0; }

void t41rx_wav_close(t41rx_wav *w) {
  if (!w) return;
  if (w->f) fclose(w->f);
  delete w;
}

}  // extern "C"

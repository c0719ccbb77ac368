/*
 * oracle/ref_shim/t41_shim.h — TEST INFRASTRUCTURE (Tier-A cross-check only).
 *
 * Minimal stand-in for the Arduino / Teensyduino environment so that the
 * reference's hot-path translation units (Process.cpp, Freq_Shift.cpp, FFT.cpp,
 * Demod.cpp, DSP_Fn.cpp, Filter.cpp, FIR.cpp, Utility.cpp, psk31.cpp) compile on
 * the host IN PLACE from /root/reference, unmodified.  Hardware classes are inert
 * stubs; the two audio queues carry real data so that ProcessIQData() sees the
 * q15 blocks the harness feeds it.  Nothing here is used by the product.
 */
#ifndef T41_REF_SHIM_H
#define T41_REF_SHIM_H

#include <type_traits>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "cmsis_port.h" /* plays the role of <arm_math.h> */

/* ---- Arduino language bits ---- */
typedef uint8_t byte;
typedef bool boolean;
typedef unsigned int uint;

#define FLASHMEM
#define DMAMEM
#define PROGMEM
#define FASTRUN
#define EXTMEM
#define F(x) x

#define HIGH 1
#define LOW 0
#define INPUT 0
#define OUTPUT 1
#define INPUT_PULLUP 2

/* Arduino's macro forms (wiring.h); the result type follows the usual arithmetic
   conversions, which is what the AGC's min(0.0, float) relies on */
#undef min
#undef max
#undef abs
#define min(a, b) ((a) < (b) ? (a) : (b))
#define max(a, b) ((a) > (b) ? (a) : (b))
#define abs(x) ((x) > 0 ? (x) : -(x))
#define constrain(amt, low, high) ((amt) < (low) ? (low) : ((amt) > (high) ? (high) : (amt)))

/* Arduino's double-precision constants; FIR.h re-defines PI & co as floats */
#define PI 3.1415926535897932384626433832795
#define HALF_PI 1.5707963267948966192313216916398
#define TWO_PI 6.283185307179586476925286766559
#define round(x) ((x) >= 0 ? (long)((x) + 0.5) : (long)((x)-0.5))

/* Arduino map() as the Teensyduino core (cores/teensy4/wiring.h; not under /root/reference, no version pinned)
   overloads it: an integral first argument does 32-bit signed long arithmetic (with the core's range-rounding
   variant), a floating-point first argument does the plain formula in that argument's own type and returns it.
   The only use on the receive path is the audio-spectrum by-product (Process.cpp:557,561,799), float argument. */
template <class T, class A, class B, class C, class D>
static inline typename std::enable_if<std::is_integral<T>::value, long>::type map(T _x, A _in_min, B _in_max, C _out_min,
                                                                                   D _out_max) {
  long x = _x, in_min = _in_min, in_max = _in_max, out_min = _out_min, out_max = _out_max;
  if ((in_max - in_min) > (out_max - out_min)) return (x - in_min) * (out_max - out_min + 1) / (in_max - in_min + 1) + out_min;
  return (x - in_min) * (out_max - out_min) / (in_max - in_min) + out_min;
}
template <class T, class A, class B, class C, class D>
static inline typename std::enable_if<std::is_floating_point<T>::value, T>::type map(T x, A in_min, B in_max, C out_min,
                                                                                      D out_max) {
  return (x - (T)in_min) * ((T)out_max - (T)out_min) / ((T)in_max - (T)in_min) + (T)out_min;
}

static inline void delay(unsigned long) {}
static inline void delayMicroseconds(unsigned long) {}
static inline unsigned long millis() { return 0; }
static inline unsigned long micros() { return 0; }
static inline void pinMode(int, int) {}
static inline void digitalWrite(int, int) {}
static inline int digitalRead(int) { return 0; }
static inline int analogRead(int) { return 0; }
static inline void analogWrite(int, int) {}
static inline int hour() { return 0; }
static inline int minute() { return 0; }
static inline int second() { return 0; }
static inline int day() { return 1; }
static inline int month() { return 1; }
static inline int year() { return 2024; }
static inline int hourFormat12() { return 12; }
static inline void setTime(time_t) {}
static inline void setSyncProvider(time_t (*)()) {}
static inline void setTime(int, int, int, int, int, int) {}
static inline char *dtostrf(double v, int w, unsigned p, char *buf) { sprintf(buf, "%*.*f", w, (int)p, v); return buf; }
static inline char *itoa(int v, char *buf, int) { sprintf(buf, "%d", v); return buf; }
static inline char *ltoa(long v, char *buf, int) { sprintf(buf, "%ld", v); return buf; }

class String {
 public:
  String() {}
  template <class T> String(T) {}
  template <class T> String(T, int) {}
  const char *c_str() const { return ""; }
  unsigned length() const { return 0; }
  template <class T> String operator+(const T &) const { return String(); }
  template <class T> String &operator+=(const T &) { return *this; }
};

class elapsedMicros {
 public:
  unsigned long v;
  elapsedMicros() : v(0) {}
  elapsedMicros(unsigned long x) : v(x) {}
  operator unsigned long() const { return v; }
  elapsedMicros &operator=(unsigned long x) { v = x; return *this; }
};
typedef elapsedMicros elapsedMillis;

class Print {
 public:
  template <class T> size_t print(T) { return 0; }
  template <class T, class U> size_t print(T, U) { return 0; }
  template <class T> size_t println(T) { return 0; }
  template <class T, class U> size_t println(T, U) { return 0; }
  size_t println() { return 0; }
  template <class... A> int printf(const char *, A...) { return 0; }
  size_t write(uint8_t) { return 1; }
  size_t write(const uint8_t *, size_t n) { return n; }
  size_t write(const char *) { return 0; }
};

class SerialStub : public Print {
 public:
  void begin(long) {}
  int available() { return 0; }
  int read() { return -1; }
  int peek() { return -1; }
  void flush() {}
  operator bool() const { return true; }
  size_t readBytes(char *, size_t) { return 0; }
  size_t readBytesUntil(char, char *, size_t) { return 0; }
  void setTimeout(long) {}
};
extern SerialStub Serial;
extern SerialStub SerialUSB1;
extern SerialStub SerialUSB2;

/* ---- Teensy Audio library ---- */
#define AUDIO_SAMPLE_RATE 44100.0f
#define AUDIO_SAMPLE_RATE_EXACT 44100.0f
#define AUDIO_BLOCK_SAMPLES 128

/* Input queue holding real q15 blocks of 128 samples (filled by the harness). */
class AudioRecordQueue {
 public:
  enum { kMaxBlocks = 64 };
  int16_t blocks[kMaxBlocks][128];
  int head, tail, count;
  AudioRecordQueue() : head(0), tail(0), count(0) {}
  void begin() {}
  void end() {}
  int available() const { return count; }
  void clear() { /* the harness never over-fills, so the >25 flush never fires */ head = tail = count = 0; }
  int16_t *readBuffer() { return count ? blocks[tail] : blocks[0]; }
  void freeBuffer() { if (count) { tail = (tail + 1) % kMaxBlocks; --count; } }
  void push(const int16_t *src) { memcpy(blocks[head], src, sizeof(blocks[0])); head = (head + 1) % kMaxBlocks; ++count; }
};

/* Output queue capturing what the chain plays. */
class AudioPlayQueue {
 public:
  int16_t last[4096];
  unsigned last_len;
  AudioPlayQueue() : last_len(0) {}
  void setMaxBuffers(int) {}
  void setBehaviour(int) {}
  int16_t *getBuffer() { return last; }
  void playBuffer() {}
  unsigned play(const int16_t *data, uint32_t len) {
    if (len > 4096) len = 4096;
    memcpy(last, data, len * sizeof(int16_t));
    last_len = len;
    return 0;
  }
  void play(int16_t) {}
};

class AudioStream {};
class AudioMixer4 { public: void gain(unsigned, float) {} };
class AudioInputI2SQuad {};
class AudioOutputI2SQuad {};
class AudioInputI2S {};
class AudioOutputI2S {};
class AudioInputUSB {};
class AudioOutputUSB {};
class AudioAmplifier { public: void gain(float) {} };
class AudioSynthWaveformSine { public: void frequency(float) {} void amplitude(float) {} void begin() {} void end() {} };
class AudioControlSGTL5000 {
 public:
  bool enable() { return true; }
  bool setAddress(int) { return true; }
  bool volume(float) { return true; }
  bool inputSelect(int) { return true; }
  bool micGain(unsigned) { return true; }
  bool lineInLevel(unsigned) { return true; }
  bool lineInLevel(unsigned, unsigned) { return true; }
  bool lineOutLevel(unsigned) { return true; }
  bool muteHeadphone() { return true; }
  bool unmuteHeadphone() { return true; }
  bool muteLineout() { return true; }
  bool unmuteLineout() { return true; }
  bool adcHighPassFilterDisable() { return true; }
  bool adcHighPassFilterEnable() { return true; }
  bool audioPreProcessorEnable() { return true; }
  bool audioPostProcessorEnable() { return true; }
  bool eqSelect(int) { return true; }
  unsigned short eqBands(float, float) { return 0; }
  unsigned short eqBands(float, float, float, float, float) { return 0; }
  bool enhanceBassEnable() { return true; }
  bool dacVolume(float) { return true; }
};
class AudioConnection { public: template <class... A> AudioConnection(A &...) {} template <class A, class B> AudioConnection(A &, int, B &, int) {} };
#define AUDIO_INPUT_LINEIN 0
#define AUDIO_INPUT_MIC 1
static inline void AudioMemory(int) {}
static inline void AudioMemory_F32(int) {}
static inline void AudioNoInterrupts() {}
static inline void AudioInterrupts() {}

/* OpenAudio_ArduinoLibrary */
class AudioEffectCompressor_F32 {
 public:
  void enableHPFilter(bool) {}
  void setThresh_dBFS(float) {}
  void setCompressionRatio(float) {}
  void setAttack_sec(float, float) {}
  void setRelease_sec(float, float) {}
  void setPreGain_dB(float) {}
  void setPreGain(float) {}
};
class AudioConvert_I16toF32 {};
class AudioConvert_F32toI16 {};
class AudioConnection_F32 { public: template <class... A> AudioConnection_F32(A &...) {} template <class A, class B> AudioConnection_F32(A &, int, B &, int) {} };

/* ---- misc hardware libraries ---- */
class Bounce { public: Bounce() {} Bounce(int, int) {} bool update() { return false; } bool read() { return true; } bool fallingEdge() { return false; } bool risingEdge() { return false; } };
class Metro { public: Metro() {} Metro(unsigned long) {} bool check() { return false; } void reset() {} void interval(unsigned long) {} };
#define DIR_NONE 0
#define DIR_CW 0x10
#define DIR_CCW 0x20
class Rotary { public: Rotary(int, int) {} void begin(bool = true) {} unsigned char process() { return 0; } };
enum si5351_clock { SI5351_CLK0, SI5351_CLK1, SI5351_CLK2 };
enum si5351_drive { SI5351_DRIVE_2MA, SI5351_DRIVE_4MA, SI5351_DRIVE_6MA, SI5351_DRIVE_8MA };
enum si5351_pll { SI5351_PLLA, SI5351_PLLB };
#define SI5351_CRYSTAL_LOAD_10PF 0
#define SI5351_CRYSTAL_LOAD_8PF 0
#define SI5351_FREQ_MULT 100ULL
class Si5351 {
 public:
  bool init(uint8_t, uint32_t, int32_t) { return true; }
  void reset() {}
  uint8_t set_freq(uint64_t, int) { return 0; }
  uint8_t set_freq_manual(uint64_t, uint64_t, int) { return 0; }
  void set_correction(int32_t, int) {}
  void drive_strength(int, int) {}
  void output_enable(int, uint8_t) {}
  void set_ms_source(int, int) {}
  void set_phase(int, uint8_t) {}
  void pll_reset(int) {}
};
class TwoWire { public: void begin() {} void setClock(long) {} void beginTransmission(int) {} int endTransmission() { return 0; } size_t write(uint8_t) { return 1; } int requestFrom(int, int) { return 0; } int read() { return 0; } int available() { return 0; } };
extern TwoWire Wire;
extern TwoWire Wire1;
class SPIClass { public: void begin() {} void setMOSI(int) {} void setSCK(int) {} void setMISO(int) {} };
extern SPIClass SPI;

/* SD-card file: backed by stdio so that the reference's WAV reader (Utility.cpp:773-888) can be run on real files */
class File : public Print {
 public:
  FILE *fp_ = 0;
  operator bool() const { return fp_ != 0; }
  int available() { return 0; }
  int read() { return -1; }
  size_t read(void *dst, size_t n) { return fp_ ? fread(dst, 1, n, fp_) : 0; }
  bool seek(uint32_t pos) { return fp_ && fseek(fp_, (long)pos, SEEK_SET) == 0; }
  uint32_t position() { return fp_ ? (uint32_t)ftell(fp_) : 0; }
  uint32_t size() {
    if (!fp_) return 0;
    const long here = ftell(fp_);
    fseek(fp_, 0, SEEK_END);
    const long n = ftell(fp_);
    fseek(fp_, here, SEEK_SET);
    return (uint32_t)n;
  }
  void close() {
    if (fp_) fclose(fp_);
    fp_ = 0;
  }
  void flush() {}
  bool isDirectory() { return false; }
  File openNextFile() { return File(); }
  const char *name() { return ""; }
  size_t readBytes(char *, size_t) { return 0; }
  size_t readBytesUntil(char, char *, size_t) { return 0; }
};
#define FILE_READ 0
#define FILE_WRITE 1
#define BUILTIN_SDCARD 254
class SDClass {
 public:
  bool begin(int = 0) { return false; }
  File open(const char *path, int mode = 0) {
    File f;
    if (mode == 0 && path) f.fp_ = fopen(path, "rb");      /* FILE_READ only */
    return f;
  }
  bool exists(const char *) { return false; }
  bool remove(const char *) { return false; }
  bool mkdir(const char *) { return false; }
};
extern SDClass SD;

class EEPROMClass {
 public:
  uint8_t read(int) { return 0; }
  void write(int, uint8_t) {}
  void update(int, uint8_t) {}
  template <class T> T &get(int, T &t) { return t; }
  template <class T> const T &put(int, const T &t) { return t; }
  int length() { return 4284; }
};
extern EEPROMClass EEPROM;

/* ---- RA8875 TFT ---- */
enum RA8875tsize { X16 = 0, X24, X32 };
enum RA8875writes { L1 = 0, L2, CGRAM, PATTERN, CURSOR };
enum RA8875boolean { LAYER = 0, TRANSPARENT, LIGHTEN, OR, AND, FLOATING };
enum RA8875sizes { RA8875_480x272, RA8875_800x480, Adafruit_480x272, Adafruit_800x480 };
#define RA8875_BLACK 0x0000
#define RA8875_WHITE 0xFFFF
#define RA8875_RED 0xF800
#define RA8875_GREEN 0x07E0
#define RA8875_BLUE 0x001F
#define RA8875_CYAN 0x07FF
#define RA8875_MAGENTA 0xF81F
#define RA8875_YELLOW 0xFFE0
#define RA8875_LIGHT_GREY 0xC618
#define RA8875_LIGHT_ORANGE 0xFD20
#define RA8875_DARK_ORANGE 0xFB60
#define RA8875_PINK 0xFCFF
#define RA8875_PURPLE 0x8017
#define RA8875_GRAYSCALE 2113
struct ILI9341_t3_font_t {};
class RA8875 : public Print {
 public:
  RA8875() {}
  RA8875(int, int) {}
  template <class... A> void begin(A...) {}
  template <class... A> void setFontScale(A...) {}
  template <class... A> void setFont(A...) {}
  void setFontDefault() {}
  template <class... A> void fillRect(A...) {}
  template <class... A> void drawRect(A...) {}
  template <class... A> void drawLine(A...) {}
  template <class... A> void drawFastVLine(A...) {}
  template <class... A> void drawFastHLine(A...) {}
  template <class... A> void drawPixel(A...) {}
  template <class... A> void drawPixels(A...) {}
  template <class... A> void drawCircle(A...) {}
  template <class... A> void fillCircle(A...) {}
  template <class... A> void fillTriangle(A...) {}
  template <class... A> void drawTriangle(A...) {}
  template <class... A> void drawRoundRect(A...) {}
  template <class... A> void fillRoundRect(A...) {}
  template <class... A> void fillWindow(A...) {}
  template <class... A> void clearScreen(A...) {}
  template <class... A> void clearMemory(A...) {}
  template <class... A> void setCursor(A...) {}
  template <class... A> void setTextColor(A...) {}
  template <class... A> void setRotation(A...) {}
  template <class... A> void writeTo(A...) {}
  template <class... A> void layerEffect(A...) {}
  template <class... A> void useLayers(A...) {}
  template <class... A> void writeRect(A...) {}
  template <class... A> void readRect(A...) {}
  template <class... A> void BTE_move(A...) {}
  template <class... A> void BTE_enable(A...) {}
  template <class... A> void setBackgroundColor(A...) {}
  template <class... A> void setForegroundColor(A...) {}
  template <class... A> void backlight(A...) {}
  template <class... A> void brightness(A...) {}
  template <class... A> void setTextWrap(A...) {}
  template <class... A> void setTextSize(A...) {}
  template <class... A> void setActiveWindow(A...) {}
  template <class... A> void setScrollWindow(A...) {}
  template <class... A> void setScrollMode(A...) {}
  template <class... A> void scroll(A...) {}
  template <class... A> void getCursor(A...) {}
  template <class... A> void drawArc(A...) {}
  template <class... A> void drawEllipse(A...) {}
  template <class... A> void fillEllipse(A...) {}
  template <class... A> void drawCurve(A...) {}
  template <class... A> void fillCurve(A...) {}
  template <class... A> void showCursor(A...) {}
  template <class... A> void setCursorBlinkRate(A...) {}
  bool busy() { return false; }
  void waitBusy(uint8_t = 0) {}
  bool waitPoll(uint8_t, uint8_t) { return true; }
  uint8_t getFontWidth() { return 8; }
  uint8_t getFontHeight() { return 16; }
  int16_t width() { return 800; }
  int16_t height() { return 480; }
  uint16_t Color565(uint8_t r, uint8_t g, uint8_t b) { return (uint16_t)(((r & 0xF8) << 8) | ((g & 0xFC) << 3) | (b >> 3)); }
  uint16_t getPixel(int16_t, int16_t) { return 0; }
  int16_t getCursorX() { return 0; }
  int16_t getCursorY() { return 0; }
};

/* Teensy core odds and ends */
#define CORE_PIN10_CONFIG (*(volatile uint32_t *)&t41_shim_scratch_reg)
extern volatile uint32_t t41_shim_scratch_reg;
extern volatile uint32_t TEMPMON_TEMPSENSE0, TEMPMON_TEMPSENSE1, CCM_ANALOG_PLL_AUDIO, CCM_ANALOG_PLL_AUDIO_NUM,
    CCM_ANALOG_PLL_AUDIO_DENOM, CCM_ANALOG_MISC2, CCM_CSCMR1, CCM_CS1CDR, CCM_CS2CDR, CCM_ANALOG_MISC1, IOMUXC_GPR_GPR1;
static inline void attachInterrupt(int, void (*)(), int) {}
static inline void detachInterrupt(int) {}
static inline int digitalPinToInterrupt(int p) { return p; }
#define CHANGE 4
#define FALLING 2
#define RISING 3
static inline void cli() {}
static inline void sei() {}
static inline void noInterrupts() {}
static inline void interrupts() {}
static inline float tempmonGetTemp() { return 40.0f; }
static inline uint32_t set_arm_clock(uint32_t f) { return f; }
extern "C" uint32_t t41_shim_f_cpu_actual;
#define F_CPU_ACTUAL t41_shim_f_cpu_actual
#define F_CPU 600000000UL
#define DEC 10
#define HEX 16
#define BIN 2
extern volatile uint32_t HW_OCOTP_ANA1;
static inline void set_audioClock(int, int32_t, uint32_t, bool = false) {}
#define CCM_CS1CDR_SAI1_CLK_PRED_MASK 0u
#define CCM_CS1CDR_SAI1_CLK_PODF_MASK 0u
#define CCM_CS1CDR_SAI1_CLK_PRED(n) 0u
#define CCM_CS1CDR_SAI1_CLK_PODF(n) 0u
#define CCM_CS2CDR_SAI2_CLK_PRED_MASK 0u
#define CCM_CS2CDR_SAI2_CLK_PODF_MASK 0u
#define CCM_CS2CDR_SAI2_CLK_PRED(n) 0u
#define CCM_CS2CDR_SAI2_CLK_PODF(n) 0u

#endif

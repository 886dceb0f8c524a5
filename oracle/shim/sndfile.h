/*
 * sndfile.h -- TEST-INFRASTRUCTURE stand-in for the subset of libsndfile the
 * reference's src/simpleaudio-sndfile.c uses (sf_open, sf_close, sf_perror,
 * sf_command, sf_readf_float/short, sf_writef_float/short).  libsndfile is an
 * un-vendored third-party dependency of the reference (configure.ac:16-88) and
 * is absent from this image.  RIFF/WAVE only: PCM16 and IEEE float32.
 * Written from the published libsndfile API names; only used to build the
 * unmodified reference CLI under oracle/_ref/ so that it can mint golden
 * vectors (tests/golden/).  Never part of the product path.
 */
#ifndef ORACLE_SHIM_SNDFILE_H
#define ORACLE_SHIM_SNDFILE_H

#include <stdint.h>

typedef struct oracle_sndfile_s SNDFILE;
typedef int64_t sf_count_t;

typedef struct SF_INFO {
    sf_count_t frames;
    int samplerate;
    int channels;
    int format;
    int sections;
    int seekable;
} SF_INFO;

enum {
    SF_FORMAT_WAV = 0x010000, SF_FORMAT_AIFF = 0x020000, SF_FORMAT_AU = 0x030000,
    SF_FORMAT_RAW = 0x040000, SF_FORMAT_PAF = 0x050000, SF_FORMAT_SVX = 0x060000,
    SF_FORMAT_NIST = 0x070000, SF_FORMAT_VOC = 0x080000, SF_FORMAT_IRCAM = 0x0A0000,
    SF_FORMAT_W64 = 0x0B0000, SF_FORMAT_MAT4 = 0x0C0000, SF_FORMAT_MAT5 = 0x0D0000,
    SF_FORMAT_PVF = 0x0E0000, SF_FORMAT_XI = 0x0F0000, SF_FORMAT_HTK = 0x100000,
    SF_FORMAT_SDS = 0x110000, SF_FORMAT_AVR = 0x120000, SF_FORMAT_WAVEX = 0x130000,
    SF_FORMAT_SD2 = 0x160000, SF_FORMAT_FLAC = 0x170000, SF_FORMAT_CAF = 0x180000,
    SF_FORMAT_WVE = 0x190000, SF_FORMAT_OGG = 0x200000, SF_FORMAT_MPC2K = 0x210000,
    SF_FORMAT_RF64 = 0x220000,
    SF_FORMAT_PCM_16 = 0x0002, SF_FORMAT_FLOAT = 0x0006,
    SF_FORMAT_SUBMASK = 0x0000FFFF, SF_FORMAT_TYPEMASK = 0x0FFF0000
};

enum { SFM_READ = 0x10, SFM_WRITE = 0x20 };
enum { SF_FALSE = 0, SF_TRUE = 1 };
enum { SFC_SET_ADD_PEAK_CHUNK = 0x1050 };

SNDFILE *sf_open(const char *path, int mode, SF_INFO *sfinfo);
int sf_close(SNDFILE *s);
int sf_perror(SNDFILE *s);
int sf_command(SNDFILE *s, int cmd, void *data, int datasize);
sf_count_t sf_readf_float(SNDFILE *s, float *ptr, sf_count_t frames);
sf_count_t sf_readf_short(SNDFILE *s, short *ptr, sf_count_t frames);
sf_count_t sf_writef_float(SNDFILE *s, const float *ptr, sf_count_t frames);
sf_count_t sf_writef_short(SNDFILE *s, const short *ptr, sf_count_t frames);

#endif

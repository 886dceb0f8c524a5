/*
 * sndfile_shim.c -- TEST-INFRASTRUCTURE WAV reader/writer behind the sndfile.h
 * stand-in (see that header).  PCM16 <-> float conversion follows libsndfile's
 * documented default: short -> float scales by 1/32768, float -> short is not
 * needed by the reference's rx/tx paths (tx writes in the format it opened).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sndfile.h"

struct oracle_sndfile_s {
    FILE *f;
    int mode;
    int is_float;	/* file sample encoding */
    int channels;
    int samplerate;
    long data_bytes;	/* written so far / remaining to read */
};

static const char *last_error = "no error";

static void put_u32(unsigned char *p, uint32_t v) { p[0]=v; p[1]=v>>8; p[2]=v>>16; p[3]=v>>24; }
static void put_u16(unsigned char *p, uint32_t v) { p[0]=v; p[1]=v>>8; }
static uint32_t get_u32(const unsigned char *p) { return p[0] | p[1]<<8 | p[2]<<16 | (uint32_t)p[3]<<24; }
static uint32_t get_u16(const unsigned char *p) { return p[0] | p[1]<<8; }

static int write_header(SNDFILE *s)
{
    unsigned char h[44];
    int bytes = s->is_float ? 4 : 2;
    memcpy(h, "RIFF", 4);
    put_u32(h + 4, 36 + (uint32_t)s->data_bytes);
    memcpy(h + 8, "WAVEfmt ", 8);
    put_u32(h + 16, 16);
    put_u16(h + 20, s->is_float ? 3 : 1);
    put_u16(h + 22, s->channels);
    put_u32(h + 24, s->samplerate);
    put_u32(h + 28, s->samplerate * s->channels * bytes);
    put_u16(h + 32, s->channels * bytes);
    put_u16(h + 34, bytes * 8);
    memcpy(h + 36, "data", 4);
    put_u32(h + 40, (uint32_t)s->data_bytes);
    if (fseek(s->f, 0, SEEK_SET) != 0)
	return -1;
    return fwrite(h, 1, 44, s->f) == 44 ? 0 : -1;
}

SNDFILE *sf_open(const char *path, int mode, SF_INFO *info)
{
    SNDFILE *s = calloc(1, sizeof(*s));
    if (!s) { last_error = "out of memory"; return NULL; }
    s->mode = mode;
    if (mode == SFM_WRITE) {
	if ((info->format & SF_FORMAT_TYPEMASK) != SF_FORMAT_WAV) {
	    last_error = "stand-in supports WAV only"; free(s); return NULL;
	}
	s->is_float = (info->format & SF_FORMAT_SUBMASK) == SF_FORMAT_FLOAT;
	s->channels = info->channels;
	s->samplerate = info->samplerate;
	s->f = fopen(path, "wb");
	if (!s->f) { last_error = "cannot open for write"; free(s); return NULL; }
	write_header(s);
	return s;
    }
    s->f = fopen(path, "rb");
    if (!s->f) { last_error = "cannot open for read"; free(s); return NULL; }
    unsigned char h[12];
    if (fread(h, 1, 12, s->f) != 12 || memcmp(h, "RIFF", 4) || memcmp(h + 8, "WAVE", 4)) {
	last_error = "not a RIFF/WAVE file"; fclose(s->f); free(s); return NULL;
    }
    int have_fmt = 0;
    for (;;) {
	unsigned char ch[8];
	if (fread(ch, 1, 8, s->f) != 8) { last_error = "no data chunk"; fclose(s->f); free(s); return NULL; }
	uint32_t len = get_u32(ch + 4);
	if (!memcmp(ch, "fmt ", 4)) {
	    unsigned char fm[40];
	    uint32_t take = len < sizeof(fm) ? len : sizeof(fm);
	    if (fread(fm, 1, take, s->f) != take) break;
	    if (len > take) fseek(s->f, len - take, SEEK_CUR);
	    int tag = get_u16(fm), bits = get_u16(fm + 14);
	    s->channels = get_u16(fm + 2);
	    s->samplerate = get_u32(fm + 4);
	    if (tag == 3 && bits == 32) s->is_float = 1;
	    else if (tag == 1 && bits == 16) s->is_float = 0;
	    else { last_error = "unsupported WAV encoding"; fclose(s->f); free(s); return NULL; }
	    have_fmt = 1;
	} else if (!memcmp(ch, "data", 4)) {
	    s->data_bytes = len;
	    break;
	} else {
	    fseek(s->f, len + (len & 1), SEEK_CUR);
	}
    }
    if (!have_fmt) { last_error = "no fmt chunk"; fclose(s->f); free(s); return NULL; }
    info->samplerate = s->samplerate;
    info->channels = s->channels;
    info->format = SF_FORMAT_WAV | (s->is_float ? SF_FORMAT_FLOAT : SF_FORMAT_PCM_16);
    info->frames = s->data_bytes / ((s->is_float ? 4 : 2) * s->channels);
    return s;
}

int sf_close(SNDFILE *s)
{
    if (!s) return -1;
    if (s->mode == SFM_WRITE)
	write_header(s);
    fclose(s->f);
    free(s);
    return 0;
}

int sf_perror(SNDFILE *s) { (void)s; fprintf(stderr, "%s\n", last_error); return 0; }
int sf_command(SNDFILE *s, int cmd, void *data, int datasize)
{ (void)s; (void)cmd; (void)data; (void)datasize; return 0; }

static sf_count_t read_raw(SNDFILE *s, void *buf, sf_count_t nsamp, int bytes)
{
    long want = (long)nsamp * bytes;
    if (want > s->data_bytes) want = s->data_bytes - s->data_bytes % bytes;
    size_t got = fread(buf, 1, want, s->f);
    s->data_bytes -= got;
    return got / bytes;
}

sf_count_t sf_readf_float(SNDFILE *s, float *ptr, sf_count_t frames)
{
    sf_count_t n = frames * s->channels;
    if (s->is_float)
	return read_raw(s, ptr, n, 4) / s->channels;
    short *tmp = malloc(sizeof(short) * (n ? n : 1));
    sf_count_t got = read_raw(s, tmp, n, 2);
    for (sf_count_t i = 0; i < got; i++)
	ptr[i] = (float)tmp[i] * (1.0f / 32768.0f);
    free(tmp);
    return got / s->channels;
}

sf_count_t sf_readf_short(SNDFILE *s, short *ptr, sf_count_t frames)
{
    sf_count_t n = frames * s->channels;
    if (!s->is_float)
	return read_raw(s, ptr, n, 2) / s->channels;
    float *tmp = malloc(sizeof(float) * (n ? n : 1));
    sf_count_t got = read_raw(s, tmp, n, 4);
    for (sf_count_t i = 0; i < got; i++) {
	float v = tmp[i] * 32768.0f;
	ptr[i] = v > 32767.f ? 32767 : v < -32768.f ? -32768 : (short)v;
    }
    free(tmp);
    return got / s->channels;
}

sf_count_t sf_writef_float(SNDFILE *s, const float *ptr, sf_count_t frames)
{
    sf_count_t n = frames * s->channels;
    if (!s->is_float) { last_error = "float write to PCM16 file unsupported"; return -1; }
    size_t w = fwrite(ptr, 4, n, s->f);
    s->data_bytes += (long)w * 4;
    return w / s->channels;
}

sf_count_t sf_writef_short(SNDFILE *s, const short *ptr, sf_count_t frames)
{
    sf_count_t n = frames * s->channels;
    if (s->is_float) { last_error = "short write to float file unsupported"; return -1; }
    size_t w = fwrite(ptr, 2, n, s->f);
    s->data_bytes += (long)w * 2;
    return w / s->channels;
}

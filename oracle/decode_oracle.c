/*
 * decode_oracle.c -- TEST INFRASTRUCTURE: the host build of the product's databits decoders.
 *
 * minimodem_b200/csrc/fsk_b200_decode_core.h is one source compiled twice: by nvcc into the
 * k_decode<KIND> kernels (the product) and here by gcc, so that the CPU tests can hold the very
 * same statements against the unmodified reference decoders in oracle/_ref/libfsk_ref.so
 * (src/databits_*.c, src/baudot.c, src/uic_codes.c) byte for byte, and the GPU tests can hold
 * the kernels against this build.  Only tests/ and __graft_entry__.smoke() load it; nothing in
 * the product does.
 */
#include <string.h>
#include "../minimodem_b200/csrc/fsk_b200_decode_core.h"

/* data words straight into decoder `kind`; reset_before[i] != 0 = a databits_decode(0,0,0,0)
 * call before word i (may be NULL); returns the bytes produced (first `cap` stored) */
unsigned int orc_decode_words(int kind, unsigned int n_data_bits, fsk_b200_decoder_state *st,
	const unsigned long long *words, unsigned int n, const unsigned char *reset_before,
	unsigned char *out, unsigned int cap)
{
    fsk_dec_sink k = { out, cap, 0 };
    for (unsigned int i = 0; i < n; i++) {
	if (reset_before && reset_before[i])
	    fsk_dec_reset(kind, st);
	fsk_dec_word(kind, n_data_bits, st, words[i], &k);
    }
    return k.n;
}

/* the frame records of one stream, as k_decode walks them (src/minimodem.c:1351, :1415-1446) */
unsigned int orc_decode_records(int kind, unsigned int shift, unsigned int n_data_bits, int msb_first,
	int do_rx_sync, unsigned long long sync_byte, fsk_b200_decoder_state *st,
	const uint32_t *records, unsigned int nrec, unsigned char *out, unsigned int cap)
{
    fsk_dec_sink k = { out, cap, 0 };
    for (unsigned int i = 0; i < nrec; i++)
	fsk_dec_record(kind, shift, n_data_bits, msb_first, do_rx_sync, sync_byte, st,
		records + 5u * i, &k);
    return k.n;
}

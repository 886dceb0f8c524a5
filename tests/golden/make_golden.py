#!/usr/bin/env python
"""Mint the golden vectors under tests/golden/ by running the UNMODIFIED
reference CLI (oracle/_ref/minimodem_ref, minimodem_ref_trace; built from
/root/reference by oracle/Makefile) on its own tests/*.test vectors
(tests/refcases.py).  Run in the build container only (it needs /root/reference):

    python tests/golden/make_golden.py

For every case it stores (tests/golden/<name>.npz):
  * stdout bytes and the stderr stat lines of the reference `--rx` run,
  * the sha256 + length of the audio samples the rx saw (float32, after the
    S16->float scaling), so that the oracle's TX restatement can be pinned,
  * the frame data words fed to the transmitter (for baudot: the output of the
    reference's own encoder),
  * one row per fsk_find_frame call (src/minimodem.c:1265,:1373): inputs and
    outputs, traced by oracle/trace_hook.c,
  * for `bins` cases: the raw complex FFT bins (b_mark, b_space) of every bit
    window the reference analysed (src/fsk.c:157-159),
  * for `audio` cases: the audio itself.
"""
import hashlib
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc          # noqa: E402
import refcases     # noqa: E402

REFTESTS = "/root/reference/tests"
REC = struct.Struct("<IIIIIf68sfIIfIIIIII")


def read_wav(path):
    b = open(path, "rb").read()
    assert b[:4] == b"RIFF" and b[8:12] == b"WAVE"
    pos, fmt = 12, None
    while pos < len(b):
        cid, ln = b[pos:pos + 4], struct.unpack("<I", b[pos + 4:pos + 8])[0]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", b[pos + 8:pos + 24])
        elif cid == b"data":
            raw = b[pos + 8:pos + 8 + ln]
            if fmt[0] == 3:
                return np.frombuffer(raw, "<f4").copy(), fmt[2], True
            return (np.frombuffer(raw, "<i2").astype(np.float32) * np.float32(1.0 / 32768.0)), fmt[2], False
        pos += 8 + ln + (ln & 1)
    raise ValueError("no data chunk")


def parse_trace(path, want_bins, want_windows):
    b = open(path, "rb").read()
    pos = 0
    calls, bins, nfft, windows = [], [], [], []
    while pos < len(b):
        (magic, frame_nsamples, try_first, try_max, try_step, limit, expect, conf, blo, bhi, ampl,
         start, fftsize, b_mark, b_space, n_fft, n_window) = REC.unpack_from(b, pos)
        assert magic == 0x46534b54
        pos += REC.size
        fb = np.frombuffer(b, "<f4", n_fft * 4, pos).reshape(n_fft, 4)
        pos += n_fft * 16
        if n_window:
            w = np.frombuffer(b, "<f4", n_window, pos)
            pos += n_window * 4
            if want_windows:
                windows.append(w.copy())
        calls.append((frame_nsamples, try_first, try_max, try_step, limit,
                      expect.split(b"\0")[0], conf, blo | (bhi << 32), ampl, start, fftsize, b_mark, b_space))
        nfft.append(n_fft)
        if want_bins:
            bins.append(fb.copy())
    return calls, nfft, (np.concatenate(bins) if bins else np.zeros((0, 4), np.float32)), windows


def run_case(case, tmp):
    text = case["text"]
    if isinstance(text, bytes):
        data = text
    else:
        data = open(os.path.join(REFTESTS, text), "rb").read()
    wav = os.path.join(tmp, "x.wav")
    trace = os.path.join(tmp, "trace.bin")
    for f in (wav, trace):
        if os.path.exists(f):
            os.unlink(f)
    subprocess.run([orc.REF_CLI, "--tx", "--file", wav] + case["tx"], input=data, check=True)
    env = dict(os.environ, ORACLE_TRACE_FILE=trace, ORACLE_TRACE_SAMPLES="0")
    r = subprocess.run([orc.REF_CLI_TRACE, "--rx", "--file", wav] + case["rx"], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    audio, rate, is_float = read_wav(wav)
    calls, nfft, bins, _ = parse_trace(trace, case["bins"], False)
    tx_mode = orc.Mode(case["mode"], **case["mkw"])
    decoder = "ascii8" if case["tx_ascii"] else tx_mode.decoder
    words = orc.ref_encode("baudot" if decoder == "baudot" else "ascii8", data)
    out = dict(
        stdout=np.frombuffer(r.stdout, np.uint8),
        stderr=np.frombuffer(r.stderr, np.uint8),
        text=np.frombuffer(data, np.uint8),
        words=words,
        audio_sha256=np.frombuffer(hashlib.sha256(audio.tobytes()).digest(), np.uint8),
        audio_len=np.array([audio.size, rate, int(is_float)], np.int64),
        call_u32=np.array([[c[0], c[1], c[2], c[3], c[9], c[10], c[11], c[12]] for c in calls], np.uint32),
        call_f32=np.array([[c[4], c[6], c[8]] for c in calls], np.float32),
        call_bits=np.array([c[7] for c in calls], np.uint64),
        call_sync=np.array([c[5] != calls[-1][5] for c in calls], np.bool_),
        expect_strings=np.array(sorted(set(c[5] for c in calls)), dtype="S68"),
        call_expect=np.array([c[5] for c in calls], dtype="S68"),
        call_nfft=np.array(nfft, np.uint32),
    )
    if case["bins"]:
        out["bins"] = bins
    if case["audio"]:
        if is_float:
            out["audio_f32"] = audio
        else:
            out["audio_s16"] = np.round(audio * 32768.0).astype(np.int16)
    return out, r


def main():
    assert orc.build_ref(), "needs /root/reference"
    only = set(sys.argv[1:])
    with tempfile.TemporaryDirectory() as tmp:
        for case in refcases.EVERY + refcases.CLI_ONLY + refcases.MORE:
            if only and case["name"] not in only:
                continue
            out, r = run_case(case, tmp)
            path = os.path.join(HERE, case["name"] + ".npz")
            np.savez_compressed(path, **out)
            err = r.stderr.decode(errors="replace").strip().splitlines()
            print("%-32s calls=%5d ffts=%6d out=%4dB %6.1fKB  %s" % (
                case["name"], len(out["call_bits"]), int(out["call_nfft"].sum()), len(r.stdout),
                os.path.getsize(path) / 1024, err[-1] if err else ""))


if __name__ == "__main__":
    main()

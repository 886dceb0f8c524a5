/* TEST-INFRASTRUCTURE config.h used to compile the unmodified reference CLI
 * into oracle/_ref/ (the reference normally generates this with autoconf,
 * configure.ac:19-76; autotools are absent here).  File-only audio. */
#define VERSION "0.24-oracle"
#define USE_ALSA 0
#define USE_PULSEAUDIO 0
#define USE_SNDIO 0
#define USE_SNDFILE 1
#define USE_BENCHMARKS 1

/*
 * lorab200.h -- C ABI of the B200-native LoRa demodulator hot path.
 *
 * This is the drop-in boundary: a Pothos::Block::work() (or any other host,
 * via ctypes / cgo / JNI) calls these entry points instead of running the
 * reference's scalar loop
 *
 *     for i < N: decd = in[i]*chirp[i](*fine[idx]); detector.feed(i, decd)   LoRaDemod.cpp:157-166
 *     value = detector.detect(power, powerAvg, fIndex, fftBuff)             LoRaDemod.cpp:172
 *         -> kissfft<float>::transform                                      LoRaDetector.hpp:32, kissfft.hh:77-157
 *         -> |bin|^2 scan, first-max argmax, dB metrics, fractional index   LoRaDetector.hpp:33-61
 *
 * for a whole batch of symbol windows in ONE fused sm_100a kernel launch.
 * All file:line citations are relative to the reference tree (myriadrf/LoRa-SDR).
 *
 * Conventions
 *   - plain C: pointers and sizes only, no C++/torch types; never throws.
 *   - every function returns LORAB200_OK (0) or a negative lorab200_status;
 *     lorab200_strerror() names it, lorab200_last_error() adds detail.
 *   - the caller owns every buffer it passes; the library never retains a
 *     caller pointer past the call (Pothos port buffers are only valid inside
 *     work(), SURVEY.md section 8b "Ownership").
 *   - a context is NOT thread-safe; distinct contexts are independent (one per
 *     Pothos block instance / per channel), each with its own CUDA stream.
 *   - complex samples are interleaved float32 pairs (re, im), i.e. the memory
 *     layout of std::complex<float> (input port 0 dtype, LoRaDemod.cpp:79).
 *   - there is NO CPU fallback: without a CUDA device lorab200_create fails
 *     with LORAB200_ERR_CUDA.
 */
#ifndef LORAB200_H
#define LORAB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LORAB200_ABI_VERSION 3 /* 2: lorab200_rx_state / lorab200_rx_packet carry the raw peak and total; lorab200_rx_host
                                * 3: lorab200_preamble_search_dev (additive) */

typedef struct lorab200_ctx lorab200_ctx;

typedef enum {
    LORAB200_OK = 0,
    LORAB200_ERR_INVALID = -1,   /* bad argument (sf out of 7..12, null pointer, ...) */
    LORAB200_ERR_CUDA = -2,      /* CUDA runtime error / no device; see lorab200_last_error */
    LORAB200_ERR_NOMEM = -3,     /* host or device allocation failed */
    LORAB200_ERR_UNSUPPORTED = -4
} lorab200_status;

/* Which dechirp table a call multiplies by. LoRaDemod keeps two and switches
 * with the frame-sync state (LoRaDemod.cpp:142,210,253). */
typedef enum {
    LORAB200_TABLE_UP = 0,    /* _upChirpTable: conj of the up-chirp   LoRaDemod.cpp:104 */
    LORAB200_TABLE_DOWN = 1,  /* _downChirpTable                       LoRaDemod.cpp:105 */
    LORAB200_TABLE_FINE = 2,  /* _fineTuneTable, N*128 entries         LoRaDemod.cpp:108-114 */
    LORAB200_TABLE_TWIDDLE = 3/* kissfft twiddles, N entries           kissfft.hh:17-22 */
} lorab200_table;

/* One result per symbol window = the outputs of LoRaDetector::detect
 * (LoRaDetector.hpp:29-64).  16 bytes, written once per window. */
typedef struct {
    int32_t index;      /* argmax bin, first maximum wins (strict >, init 0)  LoRaDetector.hpp:33-48 */
    float   power;      /* 20*log10(sqrt(max|bin|^2)) - 20*log10(N)           LoRaDetector.hpp:18,54 */
    float   power_avg;  /* 20*log10(sqrt(total - max)) - 20*log10(N)          LoRaDetector.hpp:50,53 */
    float   f_index;    /* 0.5*(right-left)/(2*peak-right-left), 0 if denom==0 LoRaDetector.hpp:56-61 */
} lorab200_result;

/* Fine-tune state of the dechirp loop, LoRaDemod.cpp:159-162.
 *   index  in : _fineTuneIndex at the first sample of window 0
 *          out: its value after the last sample of the last window
 *   step      : _finefreqError * _fineSteps, as a float (the amount the int
 *               index is decremented by per sample, in float arithmetic).
 * The recurrence advances over the N processed samples of every window, in
 * window order, regardless of `stride_samples`.  Pass NULL for the
 * TestDetector.cpp:25 form (single multiply by the chirp table, no fine table). */
typedef struct {
    int32_t index;
    float   step;
} lorab200_fine;

/* Flags for lorab200_create. */
#define LORAB200_FLAG_DEFAULT      0u
#define LORAB200_FLAG_SIMPLE_ONLY  1u  /* force the simple one-stage-per-sync kernel (debug / cross-check) */
/* The IQ buffers given to lorab200_demod_dev are never written by the kernel launched
 * immediately before on the same stream (e.g. a receiver that double-buffers its input):
 * consecutive demod launches then overlap their first loads with the previous launch's
 * tail (programmatic dependent launch).  Results are unaffected either way. */
#define LORAB200_FLAG_INPUT_STABLE 2u

/* ---- lifetime -------------------------------------------------------------- */

/* Create a demod context for spreading factor sf (7..12, N = 2^sf samples per
 * window) on CUDA device `device`.  Builds the reference's tables on the host
 * with the reference's recipes (LoRaDemod.cpp:97-114, kissfft.hh:17-22) and
 * uploads them.  Replaces: LoRaDemod::LoRaDemod + LoRaDetector ctor + kissfft
 * ctor (LoRaDemod.cpp:68-117, LoRaDetector.hpp:12-20, kissfft.hh:71-75). */
int lorab200_create(lorab200_ctx **ctx, int device, int sf, uint32_t flags);
void lorab200_destroy(lorab200_ctx *ctx);

int lorab200_sf(const lorab200_ctx *ctx);           /* spreading factor */
size_t lorab200_n(const lorab200_ctx *ctx);         /* samples per window, 2^sf */
int lorab200_device(const lorab200_ctx *ctx);

/* ---- tables ---------------------------------------------------------------- */

/* Read-only view of the host copy of a table (interleaved complex64).
 * *n_complex receives the entry count. Pointer is valid until destroy/set. */
const float *lorab200_table_host(const lorab200_ctx *ctx, int which, size_t *n_complex);

/* Device pointer of a table (for an NCCL broadcast or inspection). */
void *lorab200_table_dev(const lorab200_ctx *ctx, int which, size_t *n_complex);

/* Replace the UP or DOWN dechirp table with caller-supplied values (N complex64,
 * host memory), e.g. genChirp's down-chirp as TestDetector.cpp:14 uses. */
int lorab200_set_table(lorab200_ctx *ctx, int which, const float *host_data);

/* Re-read every table from DEVICE memory into the host mirror and the kernel's
 * per-launch constants (call after overwriting table_dev() buffers, e.g. by an
 * NCCL broadcast from rank 0). */
int lorab200_tables_sync_from_dev(lorab200_ctx *ctx);

/* Build one of the reference tables on the HOST into caller memory, without a
 * context or a GPU (the recipe lorab200_create uses): `out` receives
 * lorab200_table_len(sf, which) interleaved complex64 entries. */
size_t lorab200_table_len(int sf, int which);
int lorab200_build_table(int sf, int which, float *out);

/* Advance the fine-tune index recurrence (LoRaDemod.cpp:160-162) by n_samples steps in
 * closed form -- exactly what n_samples applications of
 *     index = int(float(index) - step); wrap once into [0, N*128)
 * give, in O(1) per wrap of the table instead of O(n_samples).  Host only, no context.
 * Returns the new index, or a negative status for bad arguments (|step| must be < N*128). */
int32_t lorab200_fine_advance(int sf, int32_t index, float step, uint64_t n_samples);

/* ChirpGenerator.hpp:21-47 genChirp<float> on the host (bit-identical recipe);
 * the signal generator the modulator and TestDetector use.  Returns the number
 * of samples generated (NN), or a negative status. */
int lorab200_gen_chirp(float *samps, int N, int ovs, int NN, float f0, int down,
                       float ampl, float *phase_accum);

/* ---- the hot path ---------------------------------------------------------- */

/*
 * Demodulate n_windows symbol windows that are already in DEVICE memory.
 *
 *   iq_dev         device pointer, complex64; window w starts at sample
 *                  w*stride_samples.  8-byte alignment is sufficient; 16-byte
 *                  aligned dense input (stride == N) takes the TMA path.
 *   table          LORAB200_TABLE_UP or LORAB200_TABLE_DOWN
 *   fine           NULL, or in/out fine-tune state (host pointer)
 *   out_dev        device pointer, n_windows lorab200_result records
 *   fft_dev        NULL or device pointer, n_windows*N complex64: the FFT
 *                  output of every window ("fft" debug port, LoRaDemod.cpp:172,324)
 *   dec_dev        NULL or device pointer, n_windows*N complex64: the dechirped
 *                  samples ("dec" debug port, LoRaDemod.cpp:164)
 *   cuda_stream    cudaStream_t to launch on; NULL is CUDA's default stream
 *                  (the usual CUDA convention, and what torch's default stream
 *                  is).  The call is asynchronous with respect to the host;
 *                  work is ordered with whatever else is on that stream.
 *
 * Replaces LoRaDemod.cpp:157-172 (and its twin :192-203) for a whole batch.
 */
int lorab200_demod_dev(lorab200_ctx *ctx, const void *iq_dev, size_t n_windows,
                       size_t stride_samples, int table, lorab200_fine *fine,
                       lorab200_result *out_dev, void *fft_dev, void *dec_dev,
                       void *cuda_stream);

/*
 * Multi-channel form: ONE launch over n_segments independent device buffers (the
 * channels of a multi-channel receiver, BASELINE config 4: 256 channels x SF10), each
 * holding windows_per_segment windows at stride_samples.  The reference would run one
 * LoRaDemod block (one thread) per channel; here the channels share a launch so that
 * small per-channel batches still fill the GPU.
 *
 *   iq_dev_list    HOST array of n_segments DEVICE pointers (complex64), n_segments
 *                  <= LORAB200_MAX_SEGMENTS; the pointers travel as a kernel parameter,
 *                  so the array may be reused as soon as the call returns
 *   out_dev        device pointer, n_segments*windows_per_segment records; the
 *                  record of window w of segment c is out_dev[c*windows_per_segment + w]
 *
 * Detector-only form (no fine-tune state, no debug dumps): per-channel fine-tune state
 * belongs to per-channel calls of lorab200_demod_dev.  Asynchronous, like demod_dev.
 */
#define LORAB200_MAX_SEGMENTS 256
int lorab200_demod_multi_dev(lorab200_ctx *ctx, const void *const *iq_dev_list, size_t n_segments,
                             size_t windows_per_segment, size_t stride_samples, int table,
                             lorab200_result *out_dev, void *cuda_stream);

/*
 * Same, with HOST buffers: copies the IQ span to the device, launches, copies
 * results (and the optional fft/dec dumps) back, and synchronises.  This is
 * the call a Pothos::Block::work() makes with its port buffers.  Pinned
 * (page-locked) host memory makes the copies asynchronous and faster but is
 * not required.
 */
int lorab200_demod_host(lorab200_ctx *ctx, const float *iq, size_t n_windows,
                        size_t stride_samples, int table, lorab200_fine *fine,
                        lorab200_result *out, float *fft, float *dec);

/* Page-locked host memory for the buffers handed to lorab200_demod_host (what a
 * block's custom BufferManager, LoRaDemod.cpp:330-358, would allocate from):
 * pinned memory lets the H2D/D2H copies run at full PCIe rate.  NULL on failure. */
void *lorab200_alloc_pinned(size_t bytes);
void lorab200_free_pinned(void *p);
/* 1 when p points into page-locked (or device) memory, 0 for ordinary pageable memory: which of the
 * two host-buffer paths a call with this pointer takes.  A block whose input BufferManager hands out
 * lorab200_alloc_pinned slabs (LoRaDemod.cpp:348-358 is where the reference picks its manager) gets
 * the direct copy. */
int lorab200_is_pinned(const void *p);

/* Pack results into the wire format of the demod's output packet: one int16
 * per symbol (LoRaDemod.cpp:260,290: _outSymbols.as<int16_t*>()[n] = int16_t(value)). */
void lorab200_pack_symbols(const lorab200_result *results, size_t n, int16_t *symbols);

/* ---- modulator ----------------------------------- */

/*
 * Batched modulator: n_packets packets of syms_per_packet symbols each (uint16, DEVICE
 * memory, as the reference encoder emits them) -> complex64 IQ in DEVICE memory, framed
 * as LoRaMod.cpp:120-229 does: 10 preamble up-chirps, the two sync-word chirps, two
 * down-chirps and a quarter down-chirp, the data symbols, max(padding, 1) symbols of zeros.
 * Packet p occupies iq_dev[p * lorab200_mod_samples(...) ...].  Replaces /lora/lora_mod
 * (LoRaMod.cpp:280-281; setSync, setPadding, setAmplitude, setOvs are the arguments).
 *
 *   mode  LORAB200_MOD_REFERENCE  genChirp's single-precision recurrences evaluated
 *                                 literally (ChirpGenerator.hpp:22-47; one thread per packet):
 *                                 sample phases bit-identical to the reference, samples
 *                                 within 2 float ulp (polar() through double cos/sin);
 *         LORAB200_MOD_IDEAL      closed-form chirps (ovs must be 1), every sample
 *                                 independent, HBM write speed; differs from the reference
 *                                 waveform by the reference's own accumulated rounding.
 *                                 iq_dev must be 16-byte aligned in this mode (two samples
 *                                 per store), else LORAB200_ERR_INVALID.
 * Asynchronous on cuda_stream (NULL = CUDA's default stream); no allocation or synchronous
 * copy in IDEAL mode (legal inside a CUDA-graph capture).
 */
#define LORAB200_MOD_REFERENCE 0
#define LORAB200_MOD_IDEAL 1
size_t lorab200_mod_samples(int sf, int ovs, size_t syms_per_packet, size_t padding); /* samples per packet */
int lorab200_mod_dev(lorab200_ctx *ctx, const uint16_t *symbols_dev, size_t n_packets,
                     size_t syms_per_packet, unsigned sync, size_t padding, float ampl, int ovs,
                     int mode, void *iq_dev, void *cuda_stream);

/* ---- receiver: sync acquisition + symbol extraction on the device (row f3) -- */

/* The members of LoRaDemod that survive a work() call (LoRaDemod.cpp:362-392).  All zero is the
 * state after activate(). */
typedef struct lorab200_rx_state {
    int32_t phase, table, sym_count, prev_value, freq_error, fine_index;
    float ffe;
    int32_t sig_error;
    float sig_power, sig_snr;
    float sig_max_mag2, sig_reserved;  /* raw peak |bin|^2 and double sum of |bin|^2 of the window the */
    double sig_total;                  /* signals were taken from (host-side, libm-exact dB values)     */
} lorab200_rx_state;

typedef struct lorab200_rx_packet {
    int32_t length;       /* symbols in the packet (the Packet's payload.length / 2); 0 = unused slot */
    int32_t error;        /* the "error" signal emitted for this packet (LoRaDemod.cpp:267) */
    float power, snr;     /* the "power" and "snr" signals (:268-269) */
    float max_mag2, reserved; /* raw detector outputs behind power / snr (LoRaDetector.hpp:42-54) */
    double total;
} lorab200_rx_packet;

/*
 * Run the reference demod block's whole work() sequence -- frame sync, sync word, down-chirps,
 * quarter chirp, data symbols (LoRaDemod.cpp:145-327) -- over n_channels independent IQ
 * streams in ONE launch, one CTA per channel, no host round trip per window.  Channel c reads
 * iq_dev + c*channel_stride_samples, n_samples samples; it stops when fewer than 2N samples
 * remain (consumed_dev[c] says how far it got: re-present the tail at the start of the next
 * call's buffer) or when max_packets packets have been written.
 *   state_dev    [n_channels] in/out; zero it (cudaMemset) for a fresh stream
 *   partial_dev  [n_channels][mtu] in/out scratch: the packet under construction between calls
 *   symbols_dev  [n_channels][max_packets][mtu] int16: packet p of channel c, info_dev[c][p].length symbols
 *   setSync / setThreshold / setMTU are the sync, thresh_db, mtu arguments.
 * Symbols, packet boundaries and the error signal are bit-identical to the reference block's;
 * power/snr use the device logarithm (within 1e-4 dB), so a squelch decision can differ only when
 * snr is within that distance of thresh_db; max_mag2 / total are the raw values behind them.
 * info_dev slots at and beyond n_packets_dev[c] are written with length 0; when max_packets packets
 * were written, more than 2N samples may remain unconsumed.  Asynchronous on cuda_stream.
 */
int lorab200_rx_dev(lorab200_ctx *ctx, const void *iq_dev, size_t n_channels, size_t channel_stride_samples,
                    size_t n_samples, unsigned sync, float thresh_db, size_t mtu,
                    lorab200_rx_state *state_dev, int16_t *partial_dev, int16_t *symbols_dev,
                    lorab200_rx_packet *info_dev, int32_t *n_packets_dev, uint64_t *consumed_dev,
                    size_t max_packets, void *cuda_stream);

/*
 * Preamble search over all candidate offsets (behavioural, not bit, parity).
 * The reference acquires a frame by hunting: every un-squelched window consumes N - value samples to
 * re-time the stream (LoRaDemod.cpp:179-237, :219), one dependent window after the other.  This call
 * demodulates the windows at offsets 0, S, 2S, ... of ONE stream (S = stride_samples; an even S takes the
 * fused kernel) against the up-chirp table in a single launch and reports, in stream order, every maximal
 * run of at least min_windows consecutive windows whose snr (power - power_avg, the reference's squelch
 * quantity, :173-174) is >= thresh_db and whose phase = (offset - index) mod N moves by at most one bin
 * from window to window: inside a train of identical up-chirps a window delta samples past a chirp
 * boundary dechirps to bin delta, so that phase is constant over a preamble whatever the offsets.
 *   found_dev    [max_found] records (device); n_found_dev receives the number of runs found, which may
 *                exceed max_found (only the first max_found, in stream order, are written)
 *   offset       first window of the run;  aligned = offset + N - index of that window: where the
 *                reference's re-timing (:219) would put the next window
 *   windows      windows in the run (measured up to 65536);  snr_db mean over the run;  power_db its peak
 * A receiver can start its state machine (lorab200_rx_dev with a zeroed state) at `aligned` instead of
 * walking there.  Uses per-context scratch (16 B + 1 B per window).  Asynchronous on cuda_stream.
 */
typedef struct lorab200_preamble {
    uint64_t offset;
    uint64_t aligned;
    int32_t phase;
    int32_t windows;
    float snr_db;
    float power_db;
} lorab200_preamble;

int lorab200_preamble_search_dev(lorab200_ctx *ctx, const void *iq_dev, size_t n_samples, size_t stride_samples,
                                 float thresh_db, int min_windows, lorab200_preamble *found_dev,
                                 int32_t *n_found_dev, uint32_t max_found, void *cuda_stream);

/*
 * The same receiver for HOST buffers -- what a block's work() calls (LoRaDemod.cpp:145-327 for the whole
 * input buffer in one call): channel c's samples are read from iq_list[c] (n_samples complex64 each,
 * pageable or page-locked), staged to the device, received by ONE launch, and the results come back
 * in the caller's host arrays (same shapes as lorab200_rx_dev's, all HOST memory; state and partial
 * are in/out and carry a stream from call to call).  power / snr of every packet (and sig_power /
 * sig_snr of the state) are recomputed from the raw values with the HOST libm: bit-identical to
 * the reference block's signals.  Synchronous: returns when the packets are there.
 * consumed[c] samples of channel c have been used; present the rest again, in front of the new
 * samples, at the next call (a Pothos block: inPort->consume(consumed)).
 */
int lorab200_rx_host(lorab200_ctx *ctx, const float *const *iq_list, size_t n_channels, size_t n_samples,
                     unsigned sync, float thresh_db, size_t mtu, lorab200_rx_state *state, int16_t *partial,
                     int16_t *symbols, lorab200_rx_packet *info, int32_t *n_packets, uint64_t *consumed,
                     size_t max_packets);

/* ---- decoder ------------------------------------- */

/* The setters of /lora/lora_decoder (LoRaDecoder.cpp:137-183); interleaving is always on (the
 * reference's non-interleaved branch only forwards the symbols), the whitening switch is
 * not consulted by the reference decoder either. */
typedef struct lorab200_decoder_cfg {
    int32_t ppm;              /* setSymbolSize, 0 = SF (must be 0 or 5..SF) */
    int32_t rdd;              /* setCodingRate "4/4".."4/8" -> 0..4 */
    int32_t explicit_header;  /* enableExplicit */
    int32_t hdr;              /* enableHdr */
    int32_t data_length;      /* setDataLength (implicit mode) */
    int32_t crcc;             /* enableCrcc */
    int32_t error_check;      /* enableErrorCheck */
} lorab200_decoder_cfg;

#define LORAB200_DECODE_DROPPED (-1) /* LoRaDecoder::drop() */
#define LORAB200_DECODE_IGNORED (-2) /* fewer than 8 symbols: no output, no drop (LoRaDecoder.cpp:208) */
#define LORAB200_DECODE_NO_ROOM (-3) /* bytes_stride too small for this packet */

/*
 * Batched decoder: n_packets demodulated symbol packets (DEVICE memory, packet p at
 * symbols_dev + p*symbols_stride, n_symbols_dev[p] valid symbols or symbols_stride each when
 * n_symbols_dev is NULL) -> payload bytes (packet p at bytes_dev + p*bytes_stride) and
 * lengths_dev[p] (>= 0, or one of the codes above).  Replaces LoRaDecoder::work()
 * (LoRaDecoder.cpp:196-397), bit-exactly, one packet per thread.  Asynchronous on cuda_stream.
 */
int lorab200_decode_dev(lorab200_ctx *ctx, const uint16_t *symbols_dev, size_t n_packets,
                        size_t symbols_stride, const int32_t *n_symbols_dev,
                        const lorab200_decoder_cfg *cfg, uint8_t *bytes_dev, size_t bytes_stride,
                        int32_t *lengths_dev, void *cuda_stream);

/* ---- encoder ---------------------------------------------------------------- */

/* The setters of /lora/lora_encoder (LoRaEncoder.cpp:104-135). */
typedef struct lorab200_encoder_cfg {
    int32_t ppm;              /* setSymbolSize, 0 = SF (must be 0 or 5..SF) */
    int32_t rdd;              /* setCodingRate "4/4".."4/8" -> 0..4 */
    int32_t explicit_header;  /* enableExplicit */
    int32_t crc;              /* enableCrc */
    int32_t whitening;        /* enableWhitening */
} lorab200_encoder_cfg;

/*
 * Batched encoder: n_packets payloads (DEVICE memory, packet p at payloads_dev + p*payload_stride,
 * lengths_dev[p] bytes or payload_stride each when lengths_dev is NULL, at most 255) -> symbols
 * (uint16, packet p at symbols_dev + p*symbols_stride) and n_symbols_dev[p]
 * (LORAB200_DECODE_NO_ROOM if symbols_stride is too small).  Replaces LoRaEncoder::work()
 * (LoRaEncoder.cpp:137-228) bit-exactly, except that the nibbles that pad the codewords up to a
 * multiple of PPM are zero here and unspecified in the reference (it reads them past the end
 * of its byte vector).  lorab200_encode_symbols() gives the symbol count for a payload length.
 */
size_t lorab200_encode_symbols(int sf, const lorab200_encoder_cfg *cfg, size_t payload_length);
int lorab200_encode_dev(lorab200_ctx *ctx, const uint8_t *payloads_dev, size_t n_packets, size_t payload_stride,
                        const int32_t *lengths_dev, const lorab200_encoder_cfg *cfg, uint16_t *symbols_dev,
                        size_t symbols_stride, int32_t *n_symbols_dev, void *cuda_stream);

/* ---- diagnostics ----------------------------------------------------------- */

/*
 * Host-buffer twins of lorab200_encode_dev / lorab200_decode_dev / lorab200_mod_dev: the same
 * arguments with every array in HOST memory (pageable is fine: a batch is a handful of packets),
 * synchronous.  They are what the /lora/lora_encoder, /lora/lora_decoder and /lora/lora_mod block
 * shells call from work() (LoRaEncoder.cpp:137-228, LoRaDecoder.cpp:196-397, LoRaMod.cpp:120-229).
 */
int lorab200_encode_host(lorab200_ctx *ctx, const uint8_t *payloads, size_t n_packets, size_t payload_stride,
                         const int32_t *lengths, const lorab200_encoder_cfg *cfg, uint16_t *symbols,
                         size_t symbols_stride, int32_t *n_symbols);
int lorab200_decode_host(lorab200_ctx *ctx, const uint16_t *symbols, size_t n_packets, size_t symbols_stride,
                         const int32_t *n_symbols, const lorab200_decoder_cfg *cfg, uint8_t *bytes,
                         size_t bytes_stride, int32_t *lengths);
int lorab200_mod_host(lorab200_ctx *ctx, const uint16_t *symbols, size_t n_packets, size_t syms_per_packet,
                      unsigned sync, size_t padding, float ampl, int ovs, int mode, float *iq);

const char *lorab200_strerror(int status);
const char *lorab200_last_error(const lorab200_ctx *ctx); /* "" if none; ctx may be NULL for create errors */
unsigned long long lorab200_kernel_launches(const lorab200_ctx *ctx); /* kernels launched by this ctx so far */
const char *lorab200_last_kernel(const lorab200_ctx *ctx);            /* name of the kernel variant last launched */
int lorab200_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LORAB200_H */

/*
 * loopback_driver.cpp -- TEST INFRASTRUCTURE ONLY (oracle/_ref/libloraloop.so).
 *
 * Runs the reference's full chain, restating TestLoopback.cpp:66-133 without the
 * Pothos runtime:
 *
 *   packets -> /lora/lora_encoder -> /lora/lora_mod -> (+ AWGN) -> IQ      (loopback_tx)
 *   IQ -> DEMOD -> /lora/lora_decoder -> packets                           (loopback_rx)
 *
 * (two halves so that both demods see the SAME samples: the reference encoder leaves the
 * padding nibbles of the last interleaver block uninitialised, so two transmit runs of
 * the same packets differ in their last few symbols)
 *
 * where encoder, modulator and decoder are the UNMODIFIED reference blocks
 * (/root/reference/LoRaEncoder.cpp, LoRaMod.cpp, LoRaDecoder.cpp compiled in place
 * against our ministream Pothos surface) and DEMOD is either the unmodified reference
 * LoRaDemod.cpp ("/lora/lora_demod") or our drop-in block shell
 * (lora-sdr_b200/host/blocks/LoRaDemodB200.cpp, registered here as
 * "/lora/lora_demod_b200").  Same bytes in, so the two demods can be compared
 * symbol for symbol and label for label.
 *
 * The reference's own test takes its noise from PothosComms' /comms/noise_source
 * (distribution unpinned, outside the tree); here it is i.i.d. N(0, sigma^2) per
 * component from std::mt19937_64 with a recorded seed.
 */
 /* The reference demod reads _finefreqError (and _prevValue) before ever writing them
 * (LoRaDemod.cpp:160,183 vs :386-392): with recycled heap memory that is a garbage fine
 * tune walk or an out-of-range table index.  The reference source is compiled here
 * unmodified, with its private members reachable so the harness can start it from the
 * zero state our block shell documents. */
#include <Pothos/Framework.hpp>
#include <cmath>
#include <complex>
#include <chrono>
#include <cstring>
#include <iostream>
#include <sstream>
#include <vector>
#define private public        /* after the standard headers: only the reference class is affected */
#include "LoRaDemod.cpp"      /* from /root/reference via -I, unmodified */
#undef private

#include <algorithm>
#include <complex>
#include <cstdint>
#include <cstring>
#include <random>
#include <string>
#include <vector>

namespace {

typedef std::complex<float> cf;

// adds seeded complex Gaussian noise to a stream (stand-in for /comms/noise_source + /comms/arithmetic ADD)
class NoiseAdder : public Pothos::Block
{
public:
    NoiseAdder(float sigma, uint64_t seed) : _sigma(sigma), _rng(seed), _dist(0.0f, 1.0f)
    {
        this->setupInput(0, typeid(cf));
        this->setupOutput(0, typeid(cf));
    }
    void work(void)
    {
        auto in = this->input(0);
        auto out = this->output(0);
        const size_t n = std::min(in->elements(), out->elements());
        if (n == 0) return;
        const cf *src = in->buffer().as<const cf *>();
        cf *dst = out->buffer().as<cf *>();
        for (size_t i = 0; i < n; i++)
        {
            if (_sigma > 0) dst[i] = src[i] + cf(_sigma * _dist(_rng), _sigma * _dist(_rng));
            else dst[i] = src[i];
        }
        in->consume(n);
        out->produce(n);
    }
private:
    float _sigma;
    std::mt19937_64 _rng;
    std::normal_distribution<float> _dist;
};

Pothos::Block *makeReferenceDemodFromZeroState(const size_t sf)
{
    LoRaDemod *d = new LoRaDemod(sf);
    d->_finefreqError = 0;
    d->_prevValue = 0;
    d->_freqError = 0;
    d->_symCount = 0;
    return d;
}
Pothos::BlockRegistry registerReferenceDemodZero("/lora/lora_demod_ref0", &makeReferenceDemodFromZeroState);

} // namespace

extern "C" {

struct LoopCfg
{
    int sf;
    int cr;            /* 0..4 = "4/4".."4/8" */
    int use_b200;      /* 0: reference LoRaDemod.cpp, 1: our block shell */
    int debug_ports;   /* b200 only: produce raw/dec/fft dumps */
    int padding;       /* mod.setPadding */
    int mtu;           /* demod.setMTU */
    int sync;          /* sync word */
    float ampl;        /* mod.setAmplitude */
    float sigma;       /* noise std-dev per component */
    float thresh;      /* demod.setThreshold */
    unsigned long long seed;
};

struct LoopOut
{
    int n_packets;                 /* decoded packets */
    int lengths[64];
    unsigned char bytes[64 * 256];
    int n_sym_packets;             /* demod output packets */
    int sym_lengths[64];
    short symbols[64 * 1024];
    unsigned long long dropped;
    unsigned long long launches;   /* kernels launched by our demod */
    unsigned long long samples;    /* samples through the demod */
    int n_labels;                  /* labels on the demod's raw port */
    char labels[256][24];
    unsigned long long label_index[256];
    int n_signals;                 /* error / power / snr emissions */
    float signal_power[64];
    int signal_error[64];
    char error[256];
    unsigned long long calls;      /* lorab200_demod_host calls made by our demod */
    double seconds;                /* wall time of the feed/run loop (blocks already constructed) */
    unsigned long long n_decoded;  /* decoded packets in total (lengths[] only keeps the first 64) */
    unsigned long long windows;    /* windows submitted to the GPU by our demod (incl. unused speculation) */
    unsigned long long pinned_inputs; /* work() calls of our demod whose input port buffer was page-locked */
};

/* Which encoder / modulator / decoder the entry points below instantiate: bit 0 = our LoRaEncoderB200,
 * bit 1 = our LoRaModB200, bit 2 = our LoRaDecoderB200 (lora-sdr_b200/host/blocks/LoRaCodecB200.cpp,
 * registered here under *_b200 paths next to the unmodified reference blocks); 0 = all reference. */
static int g_b200_codec = 0;
extern "C" void loopback_use_b200_codec(int mask) { g_b200_codec = mask; }
static const char *encoderPath(void) { return (g_b200_codec & 1) ? "/lora/lora_encoder_b200" : "/lora/lora_encoder"; }
static const char *modPath(void) { return (g_b200_codec & 2) ? "/lora/lora_mod_b200" : "/lora/lora_mod"; }
static const char *decoderPath(void) { return (g_b200_codec & 4) ? "/lora/lora_decoder_b200" : "/lora/lora_decoder"; }

/* Transmit side: packets -> encoder -> mod -> (+ noise) -> IQ samples (complex64) into iq_out. */
int loopback_tx(const LoopCfg *cfg, const unsigned char *payloads, const int *lengths, int n_packets,
                float *iq_out, unsigned long long max_samples, unsigned long long *n_samples, char *err256)
{
    try
    {
        static const char *crs[] = {"4/4", "4/5", "4/6", "4/7", "4/8"};
        std::unique_ptr<Pothos::Block> encoder(Pothos::BlockRegistry::make(encoderPath()));
        std::unique_ptr<Pothos::Block> mod(Pothos::BlockRegistry::make(modPath(), size_t(cfg->sf)));
        NoiseAdder noise(cfg->sigma, cfg->seed);
        encoder->call("setSpreadFactor", size_t(cfg->sf));
        encoder->call("setCodingRate", std::string(crs[cfg->cr]));
        mod->call("setAmplitude", cfg->ampl);
        mod->call("setPadding", size_t(cfg->padding));
        mod->call("setSync", (unsigned char)(cfg->sync));
        ministream::Topology topology;
        topology.connect(encoder.get(), "0", mod.get(), "0");
        topology.connect(mod.get(), "0", &noise, "0");
        noise.output("0")->collect = true;
        topology.commit();
        size_t ofs = 0;
        for (int p = 0; p < n_packets; p++)
        {
            Pothos::Packet pkt;
            pkt.payload = Pothos::BufferChunk(typeid(uint8_t), size_t(lengths[p]));
            std::memcpy(pkt.payload.as<void *>(), payloads + ofs, size_t(lengths[p]));
            ofs += size_t(lengths[p]);
            ministream::Topology::feedMessage(encoder.get(), "0", Pothos::Object(pkt));
        }
        topology.run();
        const auto &bytes = noise.output("0")->collectedBytes;
        const unsigned long long n = bytes.size() / sizeof(cf);
        if (n > max_samples) throw Pothos::Exception("loopback_tx", "iq_out too small");
        std::memcpy(iq_out, bytes.data(), bytes.size());
        *n_samples = n;
        return 0;
    }
    catch (const std::exception &ex)
    {
        std::strncpy(err256, ex.what(), 255);
        return -1;
    }
}

/* Encoder only: payload bytes -> symbols (uint16) of one packet, by the reference LoRaEncoder. */
int loopback_encode(const LoopCfg *cfg, const unsigned char *payload, int length,
                    unsigned short *syms_out, int max_syms, int *n_syms, char *err256)
{
    try
    {
        static const char *crs[] = {"4/4", "4/5", "4/6", "4/7", "4/8"};
        std::unique_ptr<Pothos::Block> encoder(Pothos::BlockRegistry::make(encoderPath()));
        encoder->call("setSpreadFactor", size_t(cfg->sf));
        encoder->call("setCodingRate", std::string(crs[cfg->cr]));
        ministream::Topology topology;
        encoder->output("0")->collect = true;
        topology.add(encoder.get());
        topology.commit();
        Pothos::Packet pkt;
        pkt.payload = Pothos::BufferChunk(typeid(uint8_t), size_t(length));
        std::memcpy(pkt.payload.as<void *>(), payload, size_t(length));
        ministream::Topology::feedMessage(encoder.get(), "0", Pothos::Object(pkt));
        topology.run();
        const auto &msgs = encoder->output("0")->collectedMessages;
        if (msgs.size() != 1) throw Pothos::Exception("loopback_encode", "encoder produced no packet");
        const auto &out = msgs[0].extract<Pothos::Packet>();
        const int n = int(out.payload.length / 2);
        if (n > max_syms) throw Pothos::Exception("loopback_encode", "syms_out too small");
        std::memcpy(syms_out, out.payload.as<const void *>(), size_t(n) * 2);
        *n_syms = n;
        return 0;
    }
    catch (const std::exception &ex)
    {
        std::strncpy(err256, ex.what(), 255);
        return -1;
    }
}

/* Encoder with every setter: cfg6 = sf, ppm, rdd, explicit, crc, whitening. */
int loopback_encode_cfg(const int *cfg6, const unsigned char *payload, int length,
                        unsigned short *syms_out, int max_syms, int *n_syms, char *err256)
{
    try
    {
        static const char *crs[] = {"4/4", "4/5", "4/6", "4/7", "4/8"};
        std::unique_ptr<Pothos::Block> encoder(Pothos::BlockRegistry::make(encoderPath()));
        encoder->call("setSpreadFactor", size_t(cfg6[0]));
        encoder->call("setSymbolSize", size_t(cfg6[1]));
        encoder->call("setCodingRate", std::string(crs[cfg6[2]]));
        encoder->call("enableExplicit", bool(cfg6[3] != 0));
        encoder->call("enableCrc", bool(cfg6[4] != 0));
        encoder->call("enableWhitening", bool(cfg6[5] != 0));
        ministream::Topology topology;
        encoder->output("0")->collect = true;
        topology.add(encoder.get());
        topology.commit();
        Pothos::Packet pkt;
        pkt.payload = Pothos::BufferChunk(typeid(uint8_t), size_t(length));
        std::memcpy(pkt.payload.as<void *>(), payload, size_t(length));
        ministream::Topology::feedMessage(encoder.get(), "0", Pothos::Object(pkt));
        topology.run();
        const auto &msgs = encoder->output("0")->collectedMessages;
        if (msgs.size() != 1) throw Pothos::Exception("loopback_encode_cfg", "encoder produced no packet");
        const auto &out = msgs[0].extract<Pothos::Packet>();
        const int n = int(out.payload.length / 2);
        if (n > max_syms) throw Pothos::Exception("loopback_encode_cfg", "syms_out too small");
        std::memcpy(syms_out, out.payload.as<const void *>(), size_t(n) * 2);
        *n_syms = n;
        return 0;
    }
    catch (const std::exception &ex)
    {
        std::strncpy(err256, ex.what(), 255);
        return -1;
    }
}

/* Modulator only: one packet of symbols -> IQ, by the reference LoRaMod (LoRaMod.cpp, ovs = 1). */
int loopback_mod(const LoopCfg *cfg, const unsigned short *syms, int n_syms,
                 float *iq_out, unsigned long long max_samples, unsigned long long *n_samples, char *err256)
{
    try
    {
        std::unique_ptr<Pothos::Block> mod(Pothos::BlockRegistry::make(modPath(), size_t(cfg->sf)));
        mod->call("setAmplitude", cfg->ampl);
        mod->call("setPadding", size_t(cfg->padding));
        mod->call("setSync", (unsigned char)(cfg->sync));
        if (cfg->mtu > 0) mod->call("setOvs", size_t(cfg->mtu));   /* this entry point reads cfg->mtu as the oversampling */
        ministream::Topology topology;
        mod->output("0")->collect = true;
        topology.add(mod.get());
        topology.commit();
        Pothos::Packet pkt;
        pkt.payload = Pothos::BufferChunk(typeid(uint16_t), size_t(n_syms));
        std::memcpy(pkt.payload.as<void *>(), syms, size_t(n_syms) * 2);
        ministream::Topology::feedMessage(mod.get(), "0", Pothos::Object(pkt));
        topology.run();
        const auto &bytes = mod->output("0")->collectedBytes;
        const unsigned long long n = bytes.size() / sizeof(cf);
        if (n > max_samples) throw Pothos::Exception("loopback_mod", "iq_out too small");
        std::memcpy(iq_out, bytes.data(), bytes.size());
        *n_samples = n;
        return 0;
    }
    catch (const std::exception &ex)
    {
        std::strncpy(err256, ex.what(), 255);
        return -1;
    }
}

/* Decoder only: one packet of demodulated symbols -> bytes, by the reference LoRaDecoder with every
 * setter exposed.  cfg8 = sf, ppm, rdd, explicit, hdr, data_length, crcc, error_check.
 * *length = payload length, or -1 when the decoder dropped the packet, -2 when it produced nothing. */
int loopback_decode(const int *cfg8, const unsigned short *syms, int n_syms, unsigned char *out, int out_cap,
                    int *length, char *err256)
{
    try
    {
        static const char *crs[] = {"4/4", "4/5", "4/6", "4/7", "4/8"};
        std::unique_ptr<Pothos::Block> decoder(Pothos::BlockRegistry::make(decoderPath()));
        decoder->call("setSpreadFactor", size_t(cfg8[0]));
        decoder->call("setSymbolSize", size_t(cfg8[1]));
        decoder->call("setCodingRate", std::string(crs[cfg8[2]]));
        decoder->call("enableExplicit", bool(cfg8[3] != 0));
        decoder->call("enableHdr", bool(cfg8[4] != 0));
        decoder->call("setDataLength", size_t(cfg8[5]));
        decoder->call("enableCrcc", bool(cfg8[6] != 0));
        decoder->call("enableErrorCheck", bool(cfg8[7] != 0));
        ministream::Topology topology;
        decoder->output("0")->collect = true;
        topology.add(decoder.get());
        topology.commit();
        Pothos::Packet pkt;
        pkt.payload = Pothos::BufferChunk(typeid(uint16_t), size_t(n_syms));
        std::memcpy(pkt.payload.as<void *>(), syms, size_t(n_syms) * 2);
        ministream::Topology::feedMessage(decoder.get(), "0", Pothos::Object(pkt));
        topology.run();
        const auto &msgs = decoder->output("0")->collectedMessages;
        const unsigned long long dropped = decoder->call<unsigned long long>("getDropped");
        if (msgs.empty()) { *length = dropped ? -1 : -2; return 0; }
        const auto &o = msgs[0].extract<Pothos::Packet>();
        if (int(o.payload.length) > out_cap) throw Pothos::Exception("loopback_decode", "out too small");
        std::memcpy(out, o.payload.as<const void *>(), o.payload.length);
        *length = int(o.payload.length);
        return 0;
    }
    catch (const std::exception &ex)
    {
        std::strncpy(err256, ex.what(), 255);
        return -1;
    }
}

/* Receive side: the SAME IQ samples -> DEMOD (reference or ours) -> decoder.
 * `feed` samples are appended to the demod's input per scheduler pass (0 = all at once). */
int loopback_rx(const LoopCfg *cfg, const float *iq, unsigned long long n_samples, unsigned long long feed, LoopOut *out)
{
    std::memset(out, 0, sizeof(*out));
    try
    {
        static const char *crs[] = {"4/4", "4/5", "4/6", "4/7", "4/8"};
        std::unique_ptr<Pothos::Block> demod(Pothos::BlockRegistry::make(
            cfg->use_b200 ? "/lora/lora_demod_b200" : "/lora/lora_demod_ref0", size_t(cfg->sf)));
        std::unique_ptr<Pothos::Block> decoder(Pothos::BlockRegistry::make(decoderPath()));
        decoder->call("setSpreadFactor", size_t(cfg->sf));
        decoder->call("setCodingRate", std::string(crs[cfg->cr]));
        demod->call("setSync", (unsigned char)(cfg->sync));
        demod->call("setMTU", size_t(cfg->mtu));
        demod->call("setThreshold", double(cfg->thresh));
        if (cfg->use_b200) demod->call("setDebugPorts", bool(cfg->debug_ports != 0));

        ministream::Topology topology;
        topology.connect(demod.get(), "0", decoder.get(), "0");
        demod->output("0")->collect = true;       /* also keep the symbol packets for comparison */
        topology.commit();
        const cf *src = reinterpret_cast<const cf *>(iq);
        if (feed == 0) feed = n_samples;
        /* port slabs are sized before the data flows (Pothos allocates at commit), outside the timed loop */
        demod->input("0")->_reserveBytes(size_t(feed + 4096) * sizeof(cf) * 2);
        const auto t_start = std::chrono::steady_clock::now();
        for (unsigned long long at = 0; at < n_samples; at += feed)
        {
            const unsigned long long n = std::min(feed, n_samples - at);
            ministream::Topology::feedStream(demod.get(), "0", src + at, n * sizeof(cf));
            topology.run();
        }

        out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
        out->n_decoded = decoder->output("0")->collectedMessages.size();

        /* decoded packets */
        size_t bo = 0;
        for (auto &m : decoder->output("0")->collectedMessages)
        {
            if (out->n_packets >= 64) break;
            const auto &pkt = m.extract<Pothos::Packet>();
            if (bo + pkt.payload.length > sizeof(out->bytes)) break;
            out->lengths[out->n_packets++] = int(pkt.payload.length);
            std::memcpy(out->bytes + bo, pkt.payload.as<const void *>(), pkt.payload.length);
            bo += pkt.payload.length;
        }
        /* demod symbol packets */
        size_t so = 0;
        for (auto &m : demod->output("0")->collectedMessages)
        {
            if (out->n_sym_packets >= 64) break;
            const auto &pkt = m.extract<Pothos::Packet>();
            const size_t ns = pkt.payload.length / 2;
            if (so + ns > sizeof(out->symbols) / 2) break;
            out->sym_lengths[out->n_sym_packets++] = int(ns);
            std::memcpy(out->symbols + so, pkt.payload.as<const void *>(), ns * 2);
            so += ns;
        }
        out->dropped = decoder->call<unsigned long long>("getDropped");
        if (cfg->use_b200) out->launches = demod->call<unsigned long long>("getKernelLaunches");
        if (cfg->use_b200) out->calls = demod->call<unsigned long long>("getDemodCalls");
        if (cfg->use_b200) out->windows = demod->call<unsigned long long>("getDemodWindows");
        if (cfg->use_b200) out->pinned_inputs = demod->call<unsigned long long>("getPinnedInputs");
        out->samples = demod->input("0")->totalElements();
        for (auto &l : demod->output("raw")->collectedLabels)
        {
            if (out->n_labels >= 256) break;
            std::strncpy(out->labels[out->n_labels], l.id.c_str(), 23);
            out->label_index[out->n_labels] = l.index;
            out->n_labels++;
        }
        for (auto &s : demod->signalLog)
        {
            if (s.first == "power" && out->n_signals < 64) out->signal_power[out->n_signals++] = s.second[0].convert<float>();
        }
        int ne = 0;
        for (auto &s : demod->signalLog)
            if (s.first == "error" && ne < 64) out->signal_error[ne++] = s.second[0].convert<int>();
        return 0;
    }
    catch (const std::exception &ex)
    {
        std::strncpy(out->error, ex.what(), sizeof(out->error) - 1);
        return -1;
    }
}

/* K channels through ONE LoRaDemodMultiB200 block (one launch per work() call): channel c reads
 * iq + c*n_samples.  Per channel: the symbol packets (at most 16, each at most cfg->mtu symbols), the
 * error signals in order, the samples consumed. */
int loopback_rx_multi(const LoopCfg *cfg, const float *iq, int n_channels, unsigned long long n_samples,
                      unsigned long long feed, int *n_packets, int *lengths, short *symbols, int *errors,
                      unsigned long long *consumed, unsigned long long *launches, char *err256)
{
    try
    {
        std::unique_ptr<Pothos::Block> demod(Pothos::BlockRegistry::make("/lora/lora_demod_multi_b200", size_t(cfg->sf), size_t(n_channels)));
        demod->call("setSync", (unsigned char)(cfg->sync));
        demod->call("setMTU", size_t(cfg->mtu));
        demod->call("setThreshold", double(cfg->thresh));
        ministream::Topology topology;
        topology.add(demod.get());
        for (int c = 0; c < n_channels; c++) demod->output(std::to_string(c))->collect = true;
        topology.commit();
        const cf *src = reinterpret_cast<const cf *>(iq);
        if (feed == 0) feed = n_samples;
        for (unsigned long long at = 0; at < n_samples; at += feed)
        {
            const unsigned long long n = std::min(feed, n_samples - at);
            for (int c = 0; c < n_channels; c++)
                ministream::Topology::feedStream(demod.get(), std::to_string(c), src + size_t(c) * n_samples + at, n * sizeof(cf));
            topology.run();
        }
        for (int c = 0; c < n_channels; c++)
        {
            const auto &msgs = demod->output(std::to_string(c))->collectedMessages;
            n_packets[c] = int(msgs.size());
            for (size_t p = 0; p < msgs.size() && p < 16; p++)
            {
                const auto &pkt = msgs[p].extract<Pothos::Packet>();
                const int n = int(pkt.payload.length / 2);
                lengths[c * 16 + int(p)] = n;
                std::memcpy(symbols + (size_t(c) * 16 + p) * size_t(cfg->mtu), pkt.payload.as<const void *>(), size_t(n) * 2);
            }
            consumed[c] = demod->input(std::to_string(c))->totalElements();
        }
        std::vector<int> ne(size_t(n_channels), 0);
        for (auto &s : demod->signalLog)
        {
            if (s.first != "error") continue;
            const int c = s.second[0].convert<int>();
            if (ne[size_t(c)] < 16) errors[c * 16 + ne[size_t(c)]++] = s.second[1].convert<int>();
        }
        *launches = demod->call<unsigned long long>("getKernelLaunches");
        return 0;
    }
    catch (const std::exception &ex)
    {
        std::strncpy(err256, ex.what(), 255);
        return -1;
    }
}

} /* extern "C" */

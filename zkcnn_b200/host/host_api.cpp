// C entry points of the stand-alone host side (include/zkcnn_host.h): model construction, circuit + witness build,
// one interactive proof per call.  The flow of zkh_build + zkh_prove is the reference's demo main
// (src/main_demo_vgg.cpp:20-42): construct the model, neuralNetwork::create(prover), verifier(&prover).verify().
#include "../../include/zkcnn_host.h"
#include "challenge_stream.hpp"
#include "neuralNetwork.hpp"
#include "verifier.hpp"
#include <chrono>

namespace {

thread_local std::string g_err;

class ArrayNumbers : public NumberSource {
public:
    ArrayNumbers(const double *v, uint64_t n) : v_(v, v + n), i_(0) {}
    double next() override { return i_ < v_.size() ? v_[i_++] : 0.0; }
private:
    vector<double> v_;
    size_t i_;
};

// whitespace-separated decimals, parsed with strtod in large blocks (the reference uses `ifstream >> double`)
class FastFileNumbers : public NumberSource {
public:
    explicit FastFileNumbers(const std::string &path) : f_(fopen(path.c_str(), "rb")) {
        if (!f_) throw std::runtime_error("cannot open input file " + path);
        buf_.resize(1 << 22);
    }
    ~FastFileNumbers() override { if (f_) fclose(f_); }
    double next() override { return next(nullptr); }
    double next(bool *end) {
        for (;;) {
            while (pos_ < len_ && isspace((unsigned char) buf_[pos_])) ++pos_;
            // a token must end inside the buffer (or at EOF) before it is parsed
            size_t e = pos_;
            while (e < len_ && !isspace((unsigned char) buf_[e])) ++e;
            if (pos_ < len_ && (e < len_ || eof_)) {
                char save = buf_[e];
                buf_[e] = 0;
                double x = strtod(&buf_[pos_], nullptr);
                buf_[e] = save;
                pos_ = e;
                return x;
            }
            if (eof_) { if (end) *end = true; return 0.0; }
            refill();
        }
    }
private:
    void refill() {
        size_t keep = len_ - pos_;
        memmove(&buf_[0], &buf_[pos_], keep);
        size_t got = fread(&buf_[keep], 1, buf_.size() - 1 - keep, f_);
        if (got == 0) eof_ = true;
        pos_ = 0;
        len_ = keep + got;
    }
    FILE *f_;
    std::string buf_;
    size_t pos_ = 0, len_ = 0;
    bool eof_ = false;
};

uint64_t fnv(const void *p, size_t n, uint64_t h = 0xcbf29ce484222325ULL) {
    auto *b = static_cast<const uint8_t *>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001b3ULL; }
    return h;
}

}  // namespace

struct zkh_session {
    std::unique_ptr<neuralNetwork> nn;
    prover p;
    zkcnn_b200::Transcript tr;
    vector<G> last_gens;
    bool built = false;
    int device = 0;
    FILE *table_dump = nullptr;
    vector<double> values;     // the input stream of the last zkh_input_* call (image + weights): a host rebuild for a new picture needs the weights again
    std::string input_path;
    int image_path = 0;        // path of the last zkh_set_image / zkh_prove_image
    ~zkh_session() { if (table_dump) fclose(table_dump); }
};

#define ZKH_BEGIN try {
#define ZKH_END                          \
    return 0;                            \
    } catch (const std::exception &e) {  \
        g_err = e.what();                \
        return -1;                       \
    }

extern "C" {

const char *zkh_last_error(void) { return g_err.c_str(); }

zkh_session *zkh_create(const char *model, const char *network, int pic_cnt, int device) {
    try {
        if (!model || pic_cnt < 1) throw std::invalid_argument("zkh_create: bad arguments");
        std::unique_ptr<zkh_session> s(new zkh_session);
        const std::string m = model;
        if (m == "lenet") s->nn.reset(new lenet(32, 32, 1, pic_cnt, MAX, "", "", ""));
        else if (m == "lenet_cifar") s->nn.reset(new lenetCifar(32, 32, 3, pic_cnt, MAX, "", "", ""));
        else if (m == "vgg11") s->nn.reset(new vgg11(32, 32, 3, pic_cnt, MAX, "", "", ""));
        else if (m == "vgg16") s->nn.reset(new vgg16(32, 32, 3, pic_cnt, MAX, "", "", ""));
        else if (m == "vgg") {
            if (!network || !*network) throw std::invalid_argument("zkh_create: model \"vgg\" needs a network description");
            s->nn = vgg::fromDescription(32, 3, pic_cnt, network);
        } else throw std::invalid_argument("zkh_create: unknown model " + m);
        s->device = device;
        s->p.setDevice(device);
        s->p.setTranscript(&s->tr);
        return s.release();
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

void zkh_destroy(zkh_session *s) { delete s; }

int64_t zkh_input_count(zkh_session *s) {
    try {
        if (!s) throw std::invalid_argument("null session");
        return s->nn->inputCount();
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

int zkh_input_file(zkh_session *s, const char *path) {
    ZKH_BEGIN
    if (!s || !path) throw std::invalid_argument("bad arguments");
    s->nn->setInput(std::unique_ptr<NumberSource>(new FastFileNumbers(path)));
    s->input_path = path;
    s->values.clear();
    ZKH_END
}

int64_t zkh_parse_numbers(const char *path, double *out, uint64_t cap) {
    try {
        if (!path) throw std::invalid_argument("null path");
        FastFileNumbers f(path);
        uint64_t n = 0;
        for (;; ++n) {
            bool end = false;
            const double x = f.next(&end);
            if (end) break;
            if (out && n < cap) out[n] = x;
        }
        return (int64_t) n;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

int zkh_input_values(zkh_session *s, const double *values, uint64_t n) {
    ZKH_BEGIN
    if (!s || !values) throw std::invalid_argument("bad arguments");
    if ((int64_t) n < s->nn->inputCount()) throw std::invalid_argument("zkh_input_values: too few values for this model");
    s->nn->setInput(std::unique_ptr<NumberSource>(new ArrayNumbers(values, n)));
    s->values.assign(values, values + n);
    s->input_path.clear();
    ZKH_END
}

int zkh_build(zkh_session *s) {
    ZKH_BEGIN
    if (!s) throw std::invalid_argument("null session");
    s->p.unpinWitness();
    s->nn->create(s->p, false);
    s->p.invalidateCircuit();
    s->p.pinWitness();
    s->built = true;
    ZKH_END
}

int zkh_prefetch_witness(zkh_session *s) {
    ZKH_BEGIN
    if (!s || !s->built) throw std::logic_error("zkh_prefetch_witness: call zkh_build first");
    s->p.prefetchWitness();
    ZKH_END
}

int zkh_prove(zkh_session *s, uint64_t seed, uint32_t flags, zkh_stats *out) {
    ZKH_BEGIN
    if (!s || !s->built) throw std::logic_error("zkh_prove: call zkh_build first");
    const bool os_rng = (flags & ZKH_CSPRNG_CHALLENGES) != 0;
    zkcnn_b200::ScopedChallengeStream rng(seed, os_rng ? zkcnn_b200::ChallengeStream::OS_CSPRNG
                                                : (flags & ZKH_FIAT_SHAMIR) ? zkcnn_b200::ChallengeStream::FIAT_SHAMIR : zkcnn_b200::ChallengeStream::SEEDED);
    s->tr.clear();
    rng.stream.transcript = &s->tr;
    prover &p = s->p;
    p.setWitnessResident((flags & ZKH_WITNESS_RESIDENT) != 0);
    p.setPrefetchNext((flags & ZKH_PREFETCH_NEXT) != 0);
    const double up0 = p.uploadTime(), pt0 = p.proveTime();
    const uint64_t l0 = p.gpuLaunches();
    auto t0 = std::chrono::steady_clock::now();
    verifier v(&p, p.C);
    v.checkPredicates = (flags & ZKH_PROVER_ONLY) == 0;   // full verification unless the caller asks for prover-only timing
    v.realGenerators = (flags & ZKH_REAL_GENERATORS) != 0;
    v.batchRounds = (flags & ZKH_ROUND_BY_ROUND) == 0;
    v.fiatShamir = (flags & ZKH_FIAT_SHAMIR) != 0 && !os_rng;
    v.devicePredicates = (flags & ZKH_HOST_PREDICATES) == 0;
    if ((flags & ZKH_FIXED_GENERATORS) && !s->last_gens.empty()) v.fixedGenerators = &s->last_gens;
    const bool ok = v.verify();
    auto t1 = std::chrono::steady_clock::now();
    s->last_gens = v.generators;
    if (out) {
        memset(out, 0, sizeof *out);
        out->ok = ok;
        out->n_layers = p.C.size;
        out->input_size = p.C.circuit[0].size;
        out->n_fr = s->tr.n_fr;
        out->n_g1 = s->tr.n_g1;
        out->proof_bytes = s->tr.bytes.size();
        out->fnv1a = (flags & ZKH_NO_HASH) ? 0 : s->tr.fnv1a();
        out->challenges = rng.stream.calls;
        out->gpu_launches = p.gpuLaunches() - l0;
        out->prove_s = p.proveTime() - pt0;
        out->poly_s = p.polyProverTime();
        out->upload_s = p.uploadTime() - up0;
        out->wall_s = std::chrono::duration<double>(t1 - t0).count();
        out->verifier_s = v.verifierTime() + v.verifierSlowTime() + v.polyVT;
        out->gkr_kb = p.proofSize();
        out->poly_kb = p.polyProofSize();
        out->h2d_bytes = p.lastUploadBytes();
        out->checks = ZKH_CHECKED_ROUND_SUMS | (v.checkPredicates ? ZKH_CHECKED_PREDICATES | ZKH_CHECKED_INPUT_GR | ZKH_CHECKED_G1 : 0);
    }
    ZKH_END
}

// the witness part of zkh_prove_image: returns the path taken (1 device, 2 host rebuild) or -1
static int set_image(zkh_session *s, const double *pixels, uint64_t n_pixels, uint64_t *h2d_bytes, double *seconds) {
    auto t0 = std::chrono::steady_clock::now();
    int path = 2;
    if ((int64_t) n_pixels < s->nn->imagePixels()) throw std::invalid_argument("too few pixel values");
    vector<F> image;
    if (s->nn->deviceWitnessSupported() && s->nn->quantizeImage(pixels, n_pixels, image)) {
        std::vector<uint64_t> ranges;
        const std::vector<uint8_t> want = s->nn->scaleDecisionLayers(s->p.C.size);
        s->p.generateWitnessOnDevice(image, ranges, &want);
        if (h2d_bytes) *h2d_bytes = image.size() * sizeof(F);
        if (s->nn->scalesMatch(ranges.data(), s->p.C.size)) {
            const u32 last = s->p.C.size - 1;
            s->nn->inferFromOutput(s->p.readLayer(last, s->p.C.circuit[last].size));
            path = 1;
        }
    }
    if (path == 2) {   // the circuit's structure does not fit this picture: build circuit + witness for it on the host
        if (s->values.empty() && !s->input_path.empty()) {
            FastFileNumbers f(s->input_path);
            const int64_t n = s->nn->inputCount();
            s->values.resize(n);
            for (auto &x : s->values) x = f.next();
        }
        if (s->values.empty()) throw std::logic_error("the weights are not available for a host rebuild");
        std::copy(pixels, pixels + s->nn->imagePixels(), s->values.begin());
        s->nn->setInput(std::unique_ptr<NumberSource>(new ArrayNumbers(s->values.data(), s->values.size())));
        s->p.unpinWitness();
        s->nn->create(s->p, false);
        s->p.invalidateCircuit();
        s->p.pinWitness();
        if (h2d_bytes) *h2d_bytes = 0;   // (zkh_prove uploads and counts the rebuilt witness)
    }
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return path;
}

int zkh_set_image(zkh_session *s, const double *pixels, uint64_t n_pixels) {
    if (!s || !s->built || !pixels) { g_err = "zkh_set_image: call zkh_build first"; return -1; }
    try {
        s->image_path = set_image(s, pixels, n_pixels, nullptr, nullptr);
        return s->image_path;
    } catch (const std::exception &e) {
        g_err = std::string("zkh_set_image: ") + e.what();
        return -1;
    }
}

int zkh_prove_image(zkh_session *s, const double *pixels, uint64_t n_pixels, uint64_t seed, uint32_t flags, zkh_stats *out) {
    if (!s || !s->built || !pixels) { g_err = "zkh_prove_image: call zkh_build first"; return -1; }
    int path;
    double gen_s = 0;
    uint64_t img_bytes = 0;
    try {
        path = set_image(s, pixels, n_pixels, &img_bytes, &gen_s);
    } catch (const std::exception &e) {
        g_err = std::string("zkh_prove_image: ") + e.what();
        return -1;
    }
    s->image_path = path;
    const int rc = zkh_prove(s, seed, path == 1 ? (flags | ZKH_WITNESS_RESIDENT) & ~(uint32_t) ZKH_PREFETCH_NEXT : flags & ~(uint32_t) ZKH_WITNESS_RESIDENT, out);
    if (rc == 0 && out) {
        out->witness_path = (uint32_t) path;
        out->upload_s += gen_s;
        if (path == 1) out->h2d_bytes = img_bytes;
    }
    return rc;
}

const uint8_t *zkh_proof(zkh_session *s, uint64_t *n_bytes) {
    if (!s) return nullptr;
    if (n_bytes) *n_bytes = s->tr.bytes.size();
    return s->tr.bytes.data();
}

void *zkh_context(zkh_session *s) { return s ? s->p.context() : nullptr; }

int zkh_inferred_class(zkh_session *s, int picture) {
    if (!s || picture < 0 || (size_t) picture >= s->nn->inferred.size()) return -1;
    return s->nn->inferred[picture];
}

int zkh_table_dump(zkh_session *s, const char *path) {
    ZKH_BEGIN
    if (!s) throw std::invalid_argument("null session");
    if (s->table_dump) { fclose(s->table_dump); s->table_dump = nullptr; }
    if (path) {
        s->table_dump = fopen(path, "w");
        if (!s->table_dump) throw std::runtime_error(std::string("cannot write ") + path);
    }
    s->p.setTableDump(s->table_dump);
    ZKH_END
}

int zkh_circuit_dump(zkh_session *s, const char *path, int with_hashes) {
    ZKH_BEGIN
    if (!s || !s->built || !path) throw std::invalid_argument("bad arguments");
    static const char *names[] = {"INPUT", "FFT", "IFFT", "ADD_BIAS", "RELU", "Sqr", "OPT_AVG_POOL", "MAX_POOL", "AVG_POOL",
                                  "DOT_PROD", "PADDING", "FCONN", "NCONV", "NCONV_MUL", "NCONV_ADD"};
    FILE *f = fopen(path, "w");
    if (!f) throw std::runtime_error(std::string("cannot write ") + path);
    const layeredCircuit &C = s->p.C;
    for (int i = 0; i < C.size; ++i) {
        const layer &c = C.circuit[i];
        fprintf(f, "L %d %s size %u bl %d u0 %u %d u1 %u %d v0 %u %d v1 %u %d mbu %d mbv %d ph2 %d fftbl %d zsi %u uni %zu bin %zu", i,
                names[(int) c.ty], c.size, (int) c.bit_length, c.size_u[0], (int) c.bit_length_u[0], c.size_u[1], (int) c.bit_length_u[1],
                c.size_v[0], (int) c.bit_length_v[0], c.size_v[1], (int) c.bit_length_v[1], (int) c.max_bl_u, (int) c.max_bl_v,
                (int) c.need_phase2, (int) c.fft_bit_length, c.zero_start_id, c.uni_gates.size(), c.bin_gates.size());
        if (with_hashes) {
            uint64_t hu = 0xcbf29ce484222325ULL, hb = hu, hv = hu;
            for (auto &g : c.uni_gates) { u32 t[4] = {g.g, g.u, g.lu, g.sc}; hu = fnv(t, sizeof t, hu); }
            for (auto &g : c.bin_gates) { u32 t[5] = {g.g, g.u, g.v, g.sc, g.l}; hb = fnv(t, sizeof t, hb); }
            uint64_t hou = fnv(c.ori_id_u.data(), c.ori_id_u.size() * 4), hov = fnv(c.ori_id_v.data(), c.ori_id_v.size() * 4);
            for (auto &x : s->p.val[i]) { uint8_t b[32]; x.serialize(b, 32); hv = fnv(b, 32, hv); }
            uint8_t sb[32];
            c.scale.serialize(sb, 32);
            fprintf(f, " h_uni %016lx h_bin %016lx h_oriu %016lx h_oriv %016lx h_val %016lx nval %zu h_scale %016lx", hu, hb, hou, hov, hv,
                    s->p.val[i].size(), fnv(sb, 32));
        }
        fprintf(f, "\n");
    }
    fclose(f);
    ZKH_END
}

}  // extern "C"

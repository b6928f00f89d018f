// See zk_types.hpp.  layeredCircuit::init / initSubset follow src/circuit.cpp:4-100 of the reference.
#include "zk_types.hpp"
#include "challenge_stream.hpp"
#include <stdexcept>

namespace zkcnn_b200 {

ChallengeStream *&active_challenge_stream() {
    static thread_local ChallengeStream *s = nullptr;
    return s;
}

static void os_random(uint8_t *out, size_t n) {
    static FILE *f = fopen("/dev/urandom", "rb");
    if (!f || fread(out, 1, n, f) != n) throw std::runtime_error("cannot read /dev/urandom");
}

void ChallengeStream::read(uint8_t *out, size_t n) {
    if (mode == OS_CSPRNG) { os_random(out, n); ++calls; return; }
    if (mode == FIAT_SHAMIR) {
        auto absorb = [&](const uint8_t *b, size_t len) { for (size_t i = 0; i < len; ++i) { fs_hash ^= b[i]; fs_hash *= 0x100000001b3ULL; } };
        if (transcript && transcript->bytes.size() > fs_pos) {
            absorb(transcript->bytes.data() + fs_pos, transcript->bytes.size() - fs_pos);
            fs_pos = transcript->bytes.size();
        }
        uint64_t h = fs_hash;
        const uint64_t tag[2] = {seed, calls};
        const uint8_t *tb = reinterpret_cast<const uint8_t *>(tag);
        for (size_t i = 0; i < sizeof tag; ++i) { h ^= tb[i]; h *= 0x100000001b3ULL; }
        state = h;
    }
    for (size_t i = 0; i < n; i += 8) {
        uint64_t w = next();
        size_t m = n - i < 8 ? n - i : 8;
        memcpy(out + i, &w, m);
    }
    ++calls;
}

void challenge_bytes(uint8_t *out, size_t n) {
    if (ChallengeStream *s = active_challenge_stream()) {
        s->read(out, n);
        return;
    }
    os_random(out, n);
}

ScopedChallengeStream::ScopedChallengeStream(uint64_t seed, ChallengeStream::Mode mode) : stream(seed, mode), saved(active_challenge_stream()) {
    active_challenge_stream() = &stream;
}
ScopedChallengeStream::~ScopedChallengeStream() { active_challenge_stream() = saved; }

}  // namespace zkcnn_b200

// ceil(log2 n), -1 for n == 0.  Same floating-point expression as the reference (src/utils.cpp:23-25) so that the
// layer bit lengths agree for every n (the expression is off by one at n = 2^29 and 2^31).
i8 ceilPow2BitLength(u32 n) { return n < 1e-9 ? -1 : (i8) ceil(log(n) / log(2.)); }

// two_mul[k] = 2^k for k <= q, two_mul[q + 1 + k] = -2^k  (src/circuit.cpp:90-100)
void layeredCircuit::init(u8 q_bit_size, u8 _layer_sz) {
    two_mul.assign(((size_t) q_bit_size + 1) << 1, F());
    two_mul[0] = F_ONE;
    two_mul[q_bit_size + 1] = -F_ONE;
    for (int i = 1; i <= q_bit_size; ++i) {
        two_mul[i] = two_mul[i - 1] + two_mul[i - 1];
        two_mul[i + q_bit_size + 1] = -two_mul[i];
    }
    size = _layer_sz;
    circuit.assign(size, layer());
}

// Compacts every layer's references into layer 0: operands that live in the input layer are renumbered in order of
// first use (separately for the u and the v side), ori_id_u/v keep the original positions (src/circuit.cpp:4-88).
void layeredCircuit::initSubset() {
    const u32 n0 = circuit[0].size;
    vector<u32> seen_u(n0, 0), seen_v(n0, 0);     // last layer (id) that touched the input position
    vector<u32> slot_u(n0, 0), slot_v(n0, 0);
    for (u32 i = 1; i < size; ++i) {
        layer &cur = circuit[i];
        const layer &lst = circuit[i - 1];
        bool prev_u = cur.ty == layerType::FFT || cur.ty == layerType::IFFT;
        bool prev_v = false;
        auto remap_u = [&](u32 &idx) {
            if (seen_u[idx] != i) {
                seen_u[idx] = i;
                slot_u[idx] = cur.size_u[0]++;
                cur.ori_id_u.push_back(idx);
            }
            idx = slot_u[idx];
        };
        auto remap_v = [&](u32 &idx) {
            if (seen_v[idx] != i) {
                seen_v[idx] = i;
                slot_v[idx] = cur.size_v[0]++;
                cur.ori_id_v.push_back(idx);
            }
            idx = slot_v[idx];
        };
        for (auto &g : cur.uni_gates) {
            if (!g.lu) remap_u(g.u);
            else prev_u = true;
        }
        for (auto &g : cur.bin_gates) {
            if (!g.getLayerIdU(i)) remap_u(g.u);
            else prev_u = true;
            if (!g.getLayerIdV(i)) remap_v(g.v);
            else prev_v = true;
        }
        cur.bit_length_u[0] = ceilPow2BitLength(cur.size_u[0]);
        cur.bit_length_v[0] = ceilPow2BitLength(cur.size_v[0]);
        if (prev_u) {
            if (cur.ty == layerType::FFT) {
                cur.bit_length_u[1] = cur.fft_bit_length - 1;
                cur.size_u[1] = 1u << cur.bit_length_u[1];
            } else if (cur.ty == layerType::IFFT) {
                cur.bit_length_u[1] = cur.fft_bit_length;
                cur.size_u[1] = 1u << cur.bit_length_u[1];
            } else {
                cur.size_u[1] = lst.size;
                cur.bit_length_u[1] = lst.bit_length;
            }
        } else {
            cur.size_u[1] = 0;
            cur.bit_length_u[1] = -1;
        }
        if (prev_v) {
            if (cur.ty == layerType::DOT_PROD) {
                cur.size_v[1] = lst.size >> cur.fft_bit_length;
                cur.bit_length_v[1] = lst.bit_length - cur.fft_bit_length;
            } else {
                cur.size_v[1] = lst.size;
                cur.bit_length_v[1] = lst.bit_length;
            }
        } else {
            cur.size_v[1] = 0;
            cur.bit_length_v[1] = -1;
        }
        cur.updateSize();
    }
}

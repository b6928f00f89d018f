// See neuralNetwork.hpp.  The gate order, the layout of the input layer val[0] and the quantisation rules are those of
// src/neuralNetwork.cpp of the reference (cited per function), because the order in which input-layer operands are first
// used fixes the table layout of every sumcheck (layeredCircuit::initSubset) and therefore the proof transcript.
#include "neuralNetwork.hpp"
#include <functional>
#include <thread>

using std::vector;

// ---- index helpers (src/utils.cpp:209-222) ---------------------------------------------------------------------------------
long matIdx(long x, long y, long n) { return x * n + y; }
long cubIdx(long x, long y, long z, long n, long m) { return (x * n + y) * m + z; }
long tesIdx(long w, long x, long y, long z, long n, long m, long l) { return ((w * n + x) * m + y) * l + z; }
static inline bool inside(long x, long y, long nx, long ny) { return 0 <= x && x < nx && 0 <= y && y < ny; }
static inline long sqr(long x) { return x * x; }

static void initLayer(layer &c, long size, layerType ty) {   // src/utils.cpp:193-197
    c.size = c.zero_start_id = size;
    c.bit_length = ceilPow2BitLength(size);
    c.ty = ty;
}

// getRootOfUnit (src/utils.cpp:224-232): 2^n-th root of unity; the reference obtains it by n-1 mcl square roots of -1,
// which equals ROOT32^(2^(32-n)) for the recorded getRootOfUnit(32) (tools/gen_constants.py).
F getRootOfUnit(int n) {
    if (!n) return F_ONE;
    if (n > 32) throw std::range_error("getRootOfUnit: Fr has 2-adicity 32");
    zk::fr_t w;
    memcpy(w.v, h_fr_ROOT32, 32);
    for (int i = n; i < 32; ++i) w = w.sqr();
    return F(w);
}

// in-place radix-2 NTT with the reference's twiddle convention (src/utils.cpp:105-145); flag = inverse
void fft(vector<F> &arr, int logn, bool flag) {
    const u32 len = 1u << logn;
    vector<u32> rev(len);
    vector<F> w(len);
    rev[0] = 0;
    for (u32 i = 1; i < len; ++i) rev[i] = rev[i >> 1] >> 1 | (i & 1) << (logn - 1);
    w[0] = F_ONE;
    if (len > 1) {
        w[1] = getRootOfUnit(logn);
        if (flag) F::inv(w[1], w[1]);
        for (u32 i = 2; i < len; ++i) w[i] = w[i - 1] * w[1];
    }
    for (u32 i = 0; i < len; ++i)
        if (rev[i] < i) std::swap(arr[i], arr[rev[i]]);
    for (u32 span = 2; span <= len; span <<= 1) {
        const u32 half = span >> 1, step = len / span;
        for (u32 j = 0; j < len; j += span)
            for (u32 k = 0; k < half; ++k) {
                F u = arr[j + k], v = arr[j + k + half] * w[step * k];
                arr[j + k] = u + v;
                arr[j + k + half] = u - v;
            }
    }
    if (flag) {
        F ilen;
        F::inv(ilen, F((u64) len));
        for (u32 i = 0; i < len; ++i) arr[i] = arr[i] * ilen;
    }
}

namespace {
class FileNumbers : public NumberSource {
public:
    explicit FileNumbers(const string &path) : f(path) {
        if (!f.is_open()) fprintf(stderr, "Can't find the input file!!!\n");
    }
    double next() override {
        double x = 0;
        f >> x;
        return x;
    }
private:
    std::ifstream f;
};
}  // namespace

neuralNetwork::neuralNetwork(i64 psize_x, i64 psize_y, i64 pchannel, i64 pparallel, const string &i_filename, const string &c_filename,
                             const string &o_filename)
    : pool_ty(NONE), pool_bl(0), pool_sz(0), pool_stride_bl(0), pool_stride(0), pool_layer_cnt(0), act_layer_cnt(0), conv_layer_cnt(0),
      act_ty(RELU_ACT), pic_size_x(psize_x), pic_size_y(psize_y), pic_channel(pchannel), pic_parallel(pparallel), SIZE(0),
      NCONV_FAST_SIZE(1), NCONV_SIZE(2), FFT_SIZE(5), AVE_POOL_SIZE(1), FC_SIZE(1), RELU_SIZE(1), T(0), Q_MAX(0), x_bit(0), w_bit(0),
      x_next_bit(0), o_file(o_filename) {
    (void) c_filename;   // the scale/zero-point file is opened but never read by the reference (src/neuralNetwork.cpp:32-34)
    if (!i_filename.empty()) in.reset(new FileNumbers(i_filename));
}

i64 neuralNetwork::inputCount() {
    initParam();
    return total_in_size - (pic_parallel - 1) * pic_size_x * pic_size_y * pic_channel;   // the image is read once and replicated
}

// ---- create (src/neuralNetwork.cpp:60-142) -----------------------------------------------------------------------------------
void neuralNetwork::create(prover &pr, bool only_compute) {
    if (!in) throw std::runtime_error("neuralNetwork::create: no input (file or NumberSource) set");
    initParam();
    pr.C.init(Q_BIT_SIZE, SIZE);
    pr.val.assign(SIZE, vector<F>());
    pr.aux_ops.assign(SIZE, vector<uint32_t>());
    aux_ops_ = &pr.aux_ops;
    aux_supported_ = true;
    aux_step_ = 0;
    scale_decisions_.clear();
    val = pr.val.begin();
    two_mul = pr.C.two_mul.begin();

    // the network as a flat list of stages, then one pass that emits the layers of each stage in order
    struct Stage { enum Kind { CONV, POOL, FCON } kind; size_t a, b; };
    vector<Stage> plan;
    for (size_t i = 0; i < conv_section.size(); ++i) {
        for (size_t j = 0; j < conv_section[i].size(); ++j) plan.push_back({Stage::CONV, i, j});
        if (i < pool.size()) plan.push_back({Stage::POOL, i, 0});
    }
    for (size_t i = 0; i < full_conn.size(); ++i) plan.push_back({Stage::FCON, i, 0});

    i64 layer_id = 0;
    inputLayer(pr.C.circuit[layer_id++]);
    new_nx_in = pic_size_x;
    new_ny_in = pic_size_y;
    // after a layer whose outputs get re-quantised: the scale of the next activations decides the widths of the decompositions that follow
    auto rescale = [&](i64 produced_layer) {
        x_next_bit = getNextBit((int) produced_layer);
        scale_decisions_.push_back({(int) produced_layer, x_bit, w_bit, x_next_bit});
        T = x_bit + w_bit - x_next_bit;
        Q_MAX = Q + T;
    };
    for (const Stage &st : plan) {
        auto next = [&]() -> layer & { return pr.C.circuit[layer_id]; };
        if (st.kind == Stage::CONV) {
            const convKernel &conv = conv_section[st.a][st.b];
            refreshConvParam(new_nx_in, new_ny_in, conv);
            const bool last_of_section = st.b + 1 == conv_section[st.a].size();
            pool_ty = st.a < pool.size() && last_of_section ? pool[st.a].ty : NONE;
            x_bit = x_next_bit;
            if (conv.ty == FFT) {
                paddingLayer(next(), layer_id, conv.weight_start_id);
                fftLayer(next(), layer_id);
                dotProdLayer(next(), layer_id);
                ifftLayer(next(), layer_id);
                addBiasLayer(next(), layer_id, conv.bias_start_id);
            } else if (conv.ty == NAIVE_FAST) {
                naiveConvLayer(next(), layer_id, ConvEmit::FUSED, conv.weight_start_id, conv.bias_start_id);
            } else {
                naiveConvLayer(next(), layer_id, ConvEmit::PRODUCTS, conv.weight_start_id, -1);
                naiveConvLayer(next(), layer_id, ConvEmit::SUMS, -1, conv.bias_start_id);
            }
            rescale(layer_id - 1);
            if (pool_ty != MAX) reluLayer(next(), layer_id, nx_out * ny_out * channel_out * pic_parallel);   // (max pooling clamps at zero itself)
        } else if (st.kind == Stage::POOL) {
            calcSizeAfterPool(pool[st.a]);
            if (pool[st.a].ty == AVG) avgPoolingLayer(next(), layer_id);
            else if (pool[st.a].ty == MAX) maxPoolingLayer(pr.C, layer_id, pool[st.a].dcmp_start_id, pool[st.a].max_start_id, pool[st.a].max_dcmp_start_id);
        } else {
            const fconKernel &fc = full_conn[st.a];
            pool_ty = NONE;
            refreshFCParam(fc);
            x_bit = x_next_bit;
            fullyConnLayer(next(), layer_id, fc.weight_start_id, fc.bias_start_id);
            if (st.a + 1 == full_conn.size()) break;
            rescale(layer_id - 1);
            reluLayer(next(), layer_id, channel_out * pic_parallel);
        }
    }
    if (SIZE != layer_id) throw std::logic_error("neuralNetwork::create: layer count mismatch");

    total_in_size += total_max_in_size + total_ave_in_size + total_relu_in_size;
    initLayer(pr.C.circuit[0], total_in_size, layerType::INPUT);
    if ((size_t) total_in_size != pr.val[0].size()) throw std::logic_error("neuralNetwork::create: input size mismatch");

    printInfer(pr);
    if (only_compute) return;
    pr.C.initSubset();
}

// ---- layer builders -------------------------------------------------------------------------------------------------------------
void neuralNetwork::inputLayer(layer &circuit) {   // :144-152
    initLayer(circuit, total_in_size, layerType::INPUT);
    circuit.uni_gates.reserve(total_in_size);
    for (i64 i = 0; i < total_in_size; ++i) circuit.uni_gates.emplace_back(i, 0, 0, 0);
    calcInputLayer(circuit);
}

void neuralNetwork::paddingLayer(layer &circuit, i64 &layer_id, i64 first_conv_id) {   // :154-190
    const i64 lenh = getFFTLen() >> 1;
    initLayer(circuit, lenh * channel_in * (pic_parallel + channel_out), layerType::PADDING);
    circuit.fft_bit_length = getFFTBitLen();
    const i64 lo = -padding, hx = nx_in + padding, hy = ny_in + padding;
    // activations, spatially reversed so that the FFT product is a correlation
    for (i64 p = 0; p < pic_parallel; ++p)
        for (i64 ci = 0; ci < channel_in; ++ci)
            for (i64 x = lo; x < hx; ++x)
                for (i64 y = lo; y < hy; ++y) {
                    if (!inside(x, y, nx_in, ny_in)) continue;
                    i64 g = cubIdx(p, ci, matIdx(hx - x - 1, hy - y - 1, ny_padded_in), channel_in, lenh);
                    i64 u = tesIdx(p, ci, x, y, channel_in, nx_in, ny_in);
                    circuit.uni_gates.emplace_back(g, u, layer_id - 1, 0);
                }
    // kernels, straight from the input layer
    const i64 first = pic_parallel * channel_in * lenh;
    for (i64 co = 0; co < channel_out; ++co)
        for (i64 ci = 0; ci < channel_in; ++ci)
            for (i64 x = 0; x < nx_padded_in; ++x)
                for (i64 y = 0; y < ny_padded_in; ++y) {
                    if (!inside(x, y, m, m)) continue;
                    i64 g = first + cubIdx(co, ci, matIdx(x, y, ny_padded_in), channel_in, lenh);
                    i64 u = first_conv_id + tesIdx(co, ci, x, y, channel_in, m, m);
                    circuit.uni_gates.emplace_back(g, u, 0, 0);
                }
    readConvWeight(first_conv_id);
    calcNormalLayer(circuit, layer_id);
    ++layer_id;
}

void neuralNetwork::fftLayer(layer &circuit, i64 &layer_id) {   // :192-199
    initLayer(circuit, getFFTLen() * channel_in * (pic_parallel + channel_out), layerType::FFT);
    circuit.fft_bit_length = getFFTBitLen();
    calcFFTLayer(circuit, layer_id);
    ++layer_id;
}

void neuralNetwork::dotProdLayer(layer &circuit, i64 &layer_id) {   // :201-219
    initLayer(circuit, getFFTLen() * channel_out * pic_parallel, layerType::DOT_PROD);
    circuit.need_phase2 = true;
    circuit.fft_bit_length = getFFTBitLen();
    for (i64 p = 0; p < pic_parallel; ++p)
        for (i64 co = 0; co < channel_out; ++co)
            for (i64 ci = 0; ci < channel_in; ++ci)
                circuit.bin_gates.emplace_back(matIdx(p, co, channel_out), matIdx(p, ci, channel_in), matIdx(pic_parallel + co, ci, channel_in), 0, 1);
    calcDotProdLayer(circuit, layer_id);
    ++layer_id;
}

void neuralNetwork::ifftLayer(layer &circuit, i64 &layer_id) {   // :221-230
    const i64 lenh = getFFTLen() >> 1;
    initLayer(circuit, lenh * channel_out * pic_parallel, layerType::IFFT);
    circuit.fft_bit_length = getFFTBitLen();
    F::inv(circuit.scale, F((u64) 1 << circuit.fft_bit_length));
    calcFFTLayer(circuit, layer_id);
    ++layer_id;
}

void neuralNetwork::addBiasLayer(layer &circuit, i64 &layer_id, i64 first_bias_id) {   // :232-252
    initLayer(circuit, nx_out * ny_out * channel_out * pic_parallel, layerType::ADD_BIAS);
    const i64 lenh = getFFTLen() >> 1, stride = 1 << log_stride;
    const i64 lo = -padding, hx = nx_in + padding, hy = ny_in + padding;
    for (i64 p = 0; p < pic_parallel; ++p)
        for (i64 co = 0; co < channel_out; ++co)
            for (i64 x = lo; x + m <= hx; x += stride)
                for (i64 y = lo; y + m <= hy; y += stride) {
                    i64 u = cubIdx(p, co, matIdx(hx - x - 1, hy - y - 1, ny_padded_in), channel_out, lenh);
                    i64 g = tesIdx(p, co, (x - lo) >> log_stride, (y - lo) >> log_stride, channel_out, nx_out, ny_out);
                    circuit.uni_gates.emplace_back(g, first_bias_id + co, 0, 0);
                    circuit.uni_gates.emplace_back(g, u, layer_id - 1, 0);
                }
    readBias(first_bias_id);
    calcNormalLayer(circuit, layer_id);
    ++layer_id;
}

// Direct ("naive") convolution (:254-342).  The reference has three builders over the same loop nest; here one walk over the
// (picture, output channel, input channel, output position) grid and its in-bounds kernel taps feeds the three gate shapes:
//   FUSED     one layer: out[g] = bias + sum over taps of x * w                                   (NCONV)
//   PRODUCTS  first of two layers: one gate per tap, out[k] = x * w, k counting the taps           (NCONV_MUL)
//   SUMS      second of two layers: out[g] = bias + the products of g's taps, taken in order       (NCONV_ADD)
template <class PerOutput, class PerTap> void neuralNetwork::walkConvolution(PerOutput per_output, PerTap per_tap) {
    const i64 lo = -padding, hx = nx_in + padding, hy = ny_in + padding, stride = 1 << log_stride;
    for (i64 p = 0; p < pic_parallel; ++p)
        for (i64 co = 0; co < channel_out; ++co)
            for (i64 ci = 0; ci < channel_in; ++ci)
                for (i64 x = lo; x + m <= hx; x += stride)
                    for (i64 y = lo; y + m <= hy; y += stride) {
                        const i64 g = tesIdx(p, co, (x - lo) >> log_stride, (y - lo) >> log_stride, channel_out, nx_out, ny_out);
                        per_output(g, co, ci);
                        for (i64 tx = x; tx < x + m; ++tx)
                            for (i64 ty = y; ty < y + m; ++ty)
                                if (inside(tx, ty, nx_in, ny_in))
                                    per_tap(g, tesIdx(p, ci, tx, ty, channel_in, nx_in, ny_in), tesIdx(co, ci, tx - x, ty - y, channel_in, m, m));
                    }
}

void neuralNetwork::naiveConvLayer(layer &circuit, i64 &layer_id, ConvEmit mode, i64 first_conv_id, i64 first_bias_id) {
    const u8 prev_code = 2 * (u8) (layer_id > 1);   // u from the previous layer (or the picture), v = weight in layer 0
    const bool with_bias = mode != ConvEmit::PRODUCTS && ~first_bias_id;
    i64 counter = 0;                                  // PRODUCTS: next output gate; SUMS: next product of the previous layer
    if (mode != ConvEmit::PRODUCTS) initLayer(circuit, nx_out * ny_out * channel_out * pic_parallel, mode == ConvEmit::FUSED ? layerType::NCONV : layerType::NCONV_ADD);
    if (mode == ConvEmit::FUSED) circuit.bin_gates.reserve((size_t) pic_parallel * channel_out * channel_in * nx_out * ny_out * m * m);
    walkConvolution(
        [&](i64 g, i64 co, i64 ci) { if (with_bias && ci == 0) circuit.uni_gates.emplace_back(g, first_bias_id + co, 0, 0); },
        [&](i64 g, i64 u, i64 w_idx) {
            switch (mode) {
                case ConvEmit::FUSED: circuit.bin_gates.emplace_back(g, u, first_conv_id + w_idx, 0, prev_code); break;
                case ConvEmit::PRODUCTS: circuit.bin_gates.emplace_back(counter++, u, first_conv_id + w_idx, 0, prev_code); break;
                case ConvEmit::SUMS: circuit.uni_gates.emplace_back(g, counter++, layer_id - 1, 0); break;
            }
        });
    if (mode == ConvEmit::PRODUCTS) initLayer(circuit, counter, layerType::NCONV_MUL);
    if (mode != ConvEmit::SUMS) { circuit.need_phase2 = true; readConvWeight(first_conv_id); }
    if (with_bias) readBias(first_bias_id);
    calcNormalLayer(circuit, layer_id);
    ++layer_id;
}

// ReLU (:344-439; the reference has one copy after convolutions and one after fully connected layers -- they differ only in how they
// count the activations).  For each of the B activations x of the previous layer, Q_MAX witness bits in val[0]: a sign bit and the
// magnitude bits, most significant first.  Rows of the layer:
//   [0, B)             the re-quantised positive part: (1 - sign) * (top Q-1 magnitude bits recomposed)
//   [B, 2B)            x + sign-corrected recomposition of all magnitude bits == 0            (checked to be zero: zero_start_id = B)
//   [2B, 2B + B Q_MAX) b * b - b == 0 for every witness bit
void neuralNetwork::reluLayer(layer &circuit, i64 &layer_id, i64 B) {
    const i64 n_bits = B * Q_MAX;
    const i64 first_bit = val[0].size();
    val[0].resize(val[0].size() + n_bits);
    total_relu_in_size += n_bits;
    initLayer(circuit, B * (2 + Q_MAX), layerType::RELU);
    circuit.need_phase2 = true;
    circuit.zero_start_id = B;
    const u8 prev_code = 2 * (u8) (layer_id > 1);   // binGate::l of "u in the previous layer, v in layer 0"
    auto bits_of = [&](i64 act) { return first_bit + act * Q_MAX; };

    for (i64 act = 0; act < B; ++act) {              // row act: sum_s 2^(Q-1-s) bit_s  -  sign * the same
        const i64 sign = bits_of(act);
        for (i64 s = 1; s < Q; ++s) {
            circuit.uni_gates.emplace_back(act, sign + s, 0, Q - 1 - s);
            circuit.bin_gates.emplace_back(act, sign, sign + s, Q - s + Q_BIT_SIZE, 0);
        }
    }
    for (i64 act = 0; act < B; ++act) {              // row B + act: -x + 2 sign x + sum_s 2^(Q_MAX-1-s) bit_s, and the witnesses themselves
        const i64 row = B + act, sign = bits_of(act);
        circuit.uni_gates.emplace_back(row, act, layer_id - 1, Q_BIT_SIZE + 1);
        circuit.bin_gates.emplace_back(row, act, sign, 1, prev_code);
        prepareSignBit(layer_id - 1, act, sign);
        for (i64 s = 1; s < Q_MAX; ++s) {
            circuit.uni_gates.emplace_back(row, sign + s, 0, Q_MAX - s - 1);
            prepareDecmpBit(layer_id - 1, act, sign + s, Q_MAX - s - 1);
        }
    }
    for (i64 k = 0; k < n_bits; ++k) {               // row 2B + k: bit^2 - bit
        const i64 row = 2 * B + k, bit = first_bit + k;
        circuit.bin_gates.emplace_back(row, bit, bit, 0, 0);
        circuit.uni_gates.emplace_back(row, bit, 0, Q_BIT_SIZE + 1);
    }
    calcNormalLayer(circuit, layer_id);
    ++layer_id;
}

// every pooling window of the current tensor in the reference's order (picture, channel, window row, window column): per_elem for each of
// its pool_sz x pool_sz elements (cell index, offset inside the window, index of the element in the previous layer), then per_cell once
template <class PerElem, class PerCell> void neuralNetwork::walkPooling(PerElem per_elem, PerCell per_cell) {
    for (i64 p = 0; p < pic_parallel; ++p)
        for (i64 co = 0; co < channel_out; ++co)
            for (i64 x = 0; x + pool_sz <= nx_out; x += pool_stride)
                for (i64 y = 0; y + pool_sz <= ny_out; y += pool_stride) {
                    const i64 cell = tesIdx(p, co, x >> pool_stride_bl, y >> pool_stride_bl, channel_out, new_nx_in, new_ny_in);
                    for (i64 dx = 0; dx < pool_sz; ++dx)
                        for (i64 dy = 0; dy < pool_sz; ++dy) per_elem(cell, dx, dy, tesIdx(p, co, x + dx, y + dy, channel_out, nx_out, ny_out));
                    per_cell(cell);
                }
}

// Average pooling (:441-484): cell = (window sum - its low 2 log2(pool_sz) bits) / pool_sz^2, the dropped bits being witnesses in val[0]
// that a second block of rows proves to be bits.
void neuralNetwork::avgPoolingLayer(layer &circuit, i64 &layer_id) {
    const i64 n_cells = new_nx_in * new_ny_in * channel_out * pic_parallel;
    const u8 n_low = pool_bl << 1;
    initLayer(circuit, n_cells + getPoolDecmpSize(), layerType::AVG_POOL);
    F::inv(circuit.scale, F((i64) sqr(pool_sz)));
    circuit.zero_start_id = n_cells;
    circuit.need_phase2 = true;
    const i64 first_bit = val[0].size();
    val[0].resize(val[0].size() + n_cells * n_low);
    total_ave_in_size += n_cells * n_low;
    F window_sum = F_ZERO;
    walkPooling(
        [&](i64 cell, i64, i64, i64 u) {
            circuit.uni_gates.emplace_back(cell, u, layer_id - 1, 0);
            window_sum = window_sum + val[layer_id - 1][u];
        },
        [&](i64 cell) {
            for (i64 k = 0; k < n_low; ++k) {
                const i64 slot = matIdx(cell, k, n_low), bit = first_bit + slot, row = n_cells + slot;
                circuit.uni_gates.emplace_back(cell, bit, 0, n_low - k + Q_BIT_SIZE);
                prepareFieldBit(window_sum, bit, n_low - k - 1);
                circuit.bin_gates.emplace_back(row, bit, bit, 0, 0);
                circuit.uni_gates.emplace_back(row, bit, 0, Q_BIT_SIZE + 1);
            }
            window_sum = F_ZERO;
        });
    calcNormalLayer(circuit, layer_id);
    ++layer_id;
}

// Max pooling = pool_layer_cnt layers (:486-627): differences max - x_i, then a product tree proving that one of the
// differences is zero, with range proofs (bit decompositions in val[0]) for max and for every difference.
void neuralNetwork::maxPoolingLayer(layeredCircuit &C, i64 &layer_id, i64 first_dcmp_id, i64 first_max_id, i64 first_max_dcmp_id) {
    const i64 mat_new_size = new_nx_in * new_ny_in;
    const i64 tot_new_size = mat_new_size * channel_out * pic_parallel;
    const i64 pool_sz_sqr = sqr(pool_sz);

    const i64 dcmp_cnt = getPoolDecmpSize();
    first_dcmp_id = val[0].size();
    val[0].resize(val[0].size() + dcmp_cnt);
    total_max_in_size += dcmp_cnt;
    first_max_id = val[0].size();
    val[0].resize(val[0].size() + tot_new_size);
    total_max_in_size += tot_new_size;
    const i64 max_dcmp_cnt = tot_new_size * (Q_MAX - 1);
    first_max_dcmp_id = val[0].size();
    val[0].resize(val[0].size() + max_dcmp_cnt);
    total_max_in_size += max_dcmp_cnt;

    {   // layer 0: max - x_i for every window element, and max - (bits of max) == 0
        layer &circuit = C.circuit[layer_id];
        initLayer(circuit, tot_new_size * pool_sz_sqr + tot_new_size, layerType::MAX_POOL);
        circuit.zero_start_id = tot_new_size * pool_sz_sqr;
        walkPooling(
            [&](i64 cell, i64 dx, i64 dy, i64 u_g) {
                const i64 g = cubIdx(cell, dx, dy, pool_sz, pool_sz), u_max = first_max_id + cell;
                circuit.uni_gates.emplace_back(g, u_max, 0, 0);
                circuit.uni_gates.emplace_back(g, u_g, layer_id - 1, Q_BIT_SIZE + 1);
                prepareMax(layer_id - 1, u_g, u_max);
            },
            [](i64) {});
        for (i64 i_new = 0; i_new < tot_new_size; ++i_new) {
            const i64 g_new = circuit.zero_start_id + i_new, u_new = first_max_id + i_new;
            circuit.uni_gates.emplace_back(g_new, u_new, 0, Q_BIT_SIZE + 1);
            for (i64 bit = 0; bit < Q_MAX - 1; ++bit) {
                const i64 u_bit = first_max_dcmp_id + matIdx(i_new, bit, Q_MAX - 1);
                circuit.uni_gates.emplace_back(g_new, u_bit, 0, Q_MAX - 2 - bit);
                prepareDecmpBit(0, u_new, u_bit, Q_MAX - 2 - bit);
            }
        }
        calcNormalLayer(circuit, layer_id);
        ++layer_id;
    }

    i64 contain_max_ly = 1, ksize = pool_sz_sqr;
    while (!(ksize & 1)) { ksize >>= 1; ++contain_max_ly; }
    ksize = pool_sz_sqr;
    for (int i = 1; i < pool_layer_cnt; ++i) {
        layer &circuit = C.circuit[layer_id];
        const bool last = i == pool_layer_cnt - 1;
        const i64 size = tot_new_size * (((ksize + 1) >> 1) + (i64) (i == 1) * ksize) + (i64) last * tot_new_size * Q_MAX +
                         (i64) last * tot_new_size * pool_sz_sqr * (Q_MAX - 1);
        initLayer(circuit, size, layerType::MAX_POOL);
        circuit.need_phase2 = true;

        i64 before_mul = 0;
        if (last) {   // the pooled tensor itself, recomposed from the top Q-1 bits of max
            before_mul = tot_new_size;
            for (i64 g = 0; g < tot_new_size; ++g)
                for (i64 j = 0; j < Q - 1; ++j)
                    circuit.uni_gates.emplace_back(g, first_max_dcmp_id + matIdx(g, j, Q_MAX - 1), 0, Q - 2 - j);
        }
        // one level of the product tree over the differences
        const i64 half = (ksize + 1) >> 1;
        for (i64 cnt = 0; cnt < tot_new_size; ++cnt) {
            const i64 v_max = first_max_id + cnt;
            for (i64 j = 0; (j << 1) < ksize; ++j) {
                const i64 g = before_mul + matIdx(cnt, j, half);
                const i64 u = matIdx(cnt, j << 1, ksize);
                if ((j << 1 | 1) < ksize) circuit.bin_gates.emplace_back(g, u, matIdx(cnt, j << 1 | 1, ksize), 0, layer_id > 1);
                else if (i == contain_max_ly) circuit.bin_gates.emplace_back(g, u, v_max, 0, 2 * (u8) (layer_id > 1));
                else circuit.uni_gates.emplace_back(g, u, layer_id - 1, 0);
            }
        }
        if (i == 1) {   // range proof of every difference
            const i64 minus_cnt = tot_new_size * ksize;
            const i64 minus_new_cnt = tot_new_size * half;
            circuit.zero_start_id = minus_new_cnt;
            for (i64 v = 0; v < minus_cnt; ++v) {
                const i64 g = minus_new_cnt + v;
                circuit.uni_gates.emplace_back(g, v, layer_id - 1, Q_BIT_SIZE + 1);
                for (i64 bit = 0; bit < Q_MAX - 1; ++bit) {
                    const i64 u = first_dcmp_id + matIdx(v, bit, Q_MAX - 1);
                    circuit.uni_gates.emplace_back(g, u, 0, Q_MAX - 2 - bit);
                    prepareDecmpBit(layer_id - 1, v, u, Q_MAX - 2 - bit);
                }
            }
        } else if (last) {   // every bit of every difference is a bit
            const i64 minus_cnt = tot_new_size * pool_sz_sqr;
            circuit.zero_start_id = before_mul;
            for (i64 j = 0; j < minus_cnt; ++j) {
                const i64 g = before_mul + tot_new_size + j, u = first_dcmp_id + j;
                circuit.bin_gates.emplace_back(g, u, u, 0, 0);
                circuit.uni_gates.emplace_back(g, u, 0, Q_BIT_SIZE + 1);
            }
        }
        ksize = half;
        calcNormalLayer(circuit, layer_id);
        ++layer_id;
    }
}

void neuralNetwork::fullyConnLayer(layer &circuit, i64 &layer_id, i64 first_fc_id, i64 first_bias_id) {   // :629-649
    initLayer(circuit, channel_out * pic_parallel, layerType::FCONN);
    circuit.need_phase2 = true;
    const u8 l_code = 2 * (u8) (layer_id > 1);
    for (i64 p = 0; p < pic_parallel; ++p)
        for (i64 co = 0; co < channel_out; ++co) {
            const i64 g = matIdx(p, co, channel_out);
            circuit.uni_gates.emplace_back(g, first_bias_id + co, 0, 0);
            for (i64 ci = 0; ci < channel_in; ++ci)
                circuit.bin_gates.emplace_back(g, matIdx(p, ci, channel_in), first_fc_id + matIdx(co, ci, channel_in), 0, l_code);
        }
    readFconWeight(first_fc_id);
    readBias(first_bias_id);
    calcNormalLayer(circuit, layer_id);
    ++layer_id;
}

// ---- shape bookkeeping -------------------------------------------------------------------------------------------------------------
void neuralNetwork::refreshConvParam(i64 new_nx, i64 new_ny, const convKernel &conv) {   // :651-670
    nx_in = new_nx;
    ny_in = new_ny;
    padding = conv.padding;
    nx_padded_in = nx_in + conv.padding * 2;
    ny_padded_in = ny_in + conv.padding * 2;
    m = conv.size;
    channel_in = conv.channel_in;
    channel_out = conv.channel_out;
    log_stride = conv.stride_bl;
    nx_out = ((nx_padded_in - m) >> log_stride) + 1;
    ny_out = ((ny_padded_in - m) >> log_stride) + 1;
    new_nx_in = nx_out;
    new_ny_in = ny_out;
    conv_layer_cnt = conv.ty == FFT ? FFT_SIZE : conv.ty == NAIVE ? NCONV_SIZE : NCONV_FAST_SIZE;
}

void neuralNetwork::refreshFCParam(const fconKernel &fc) {   // :672-677
    nx_in = nx_out = m = 1;
    ny_in = ny_out = 1;
    channel_in = fc.channel_in;
    channel_out = fc.channel_out;
}

i64 neuralNetwork::getFFTLen() const { return 1L << getFFTBitLen(); }
i8 neuralNetwork::getFFTBitLen() const { return ceilPow2BitLength((u32) nx_padded_in * ny_padded_in) + 1; }   // :683-685

i64 neuralNetwork::getPoolDecmpSize() const {   // :786-793
    if (pool_ty == AVG) return new_nx_in * new_ny_in * (pool_bl << 1) * channel_out * pic_parallel;
    if (pool_ty == MAX) return new_nx_in * new_ny_in * sqr(pool_sz) * channel_out * pic_parallel * (Q_MAX - 1);
    throw std::logic_error("getPoolDecmpSize without pooling");
}

void neuralNetwork::calcSizeAfterPool(const poolKernel &p) {   // :795-803
    pool_sz = p.size;
    pool_bl = ceilPow2BitLength(pool_sz);
    pool_stride_bl = p.stride_bl;
    pool_stride = 1 << p.stride_bl;
    pool_layer_cnt = p.ty == MAX ? 1 + ceilPow2BitLength(sqr(p.size) + 1) : AVE_POOL_SIZE;
    new_nx_in = ((nx_out - pool_sz) >> pool_stride_bl) + 1;
    new_ny_in = ((ny_out - pool_sz) >> pool_stride_bl) + 1;
}

// layout of val[0]: [image x pic_parallel][per conv: kernel, bias][per fc: kernel, bias] then, appended while the circuit is
// built, the bit-decomposition witnesses of ReLU / pooling layers (:687-750)
void neuralNetwork::initParam() {
    act_layer_cnt = RELU_SIZE;
    i64 total_conv_layer_cnt = 0, total_pool_layer_cnt = 0;
    total_in_size = total_para_size = total_relu_in_size = total_ave_in_size = total_max_in_size = 0;
    i64 pos = pic_size_x * pic_size_y * pic_channel * pic_parallel;
    new_nx_in = pic_size_x;
    new_ny_in = pic_size_y;
    for (size_t i = 0; i < conv_section.size(); ++i) {
        auto &sec = conv_section[i];
        for (auto &conv : sec) {
            refreshConvParam(new_nx_in, new_ny_in, conv);
            conv.weight_start_id = pos;
            const u32 para_size = sqr(m) * channel_in * channel_out;
            pos += para_size;
            total_para_size += para_size;
            conv.bias_start_id = pos;
            pos += channel_out;
            total_para_size += channel_out;
        }
        total_conv_layer_cnt += sec.size() * (conv_layer_cnt + act_layer_cnt);
        if (i >= pool.size()) continue;
        calcSizeAfterPool(pool[i]);
        total_pool_layer_cnt += pool_layer_cnt;
        if (pool[i].ty == MAX && act_ty == RELU_ACT) total_conv_layer_cnt -= act_layer_cnt;
    }
    for (size_t i = 0; i < full_conn.size(); ++i) {
        auto &fc = full_conn[i];
        refreshFCParam(fc);
        fc.weight_start_id = pos;
        const u32 para_size = channel_out * channel_in;
        pos += para_size;
        total_para_size += para_size;
        fc.bias_start_id = pos;
        pos += channel_out;
        total_para_size += channel_out;
    }
    total_in_size = pos;
    SIZE = 1 + total_conv_layer_cnt + total_pool_layer_cnt + (FC_SIZE + RELU_SIZE) * full_conn.size();
    if (!full_conn.empty()) SIZE -= RELU_SIZE;
}

// ---- inputs and quantisation (:805-897) ----------------------------------------------------------------------------------------------
// scale bits: the largest i with (mx - mn) * 2^i <= 2^(Q-1) - 1
static int quantBits(double mx, double mn, i64 Q) {
    int b = (int) (log(((1 << (Q - 1)) - 1) / (mx - mn)) / log(2));
    if ((int) ((mx - mn) * exp2(b)) > (1 << (Q - 1)) - 1) --b;
    return b;
}

void neuralNetwork::calcInputLayer(layer &circuit) {
    val[0].resize(circuit.size);
    const i64 n_pix = pic_channel * pic_size_x * pic_size_y;
    vector<double> dat(n_pix);
    double mx = -10000, mn = 10000;
    for (auto &x : dat) {
        x = in->next();
        mx = max(mx, x);
        mn = min(mn, x);
    }
    x_next_bit = quantBits(mx, mn, Q);
    image_bit_ = x_next_bit;
    auto it = val[0].begin();
    for (i64 p = 0; p < pic_parallel; ++p)
        for (i64 i = 0; i < n_pix; ++i) *it++ = F((i64) (dat[i] * exp2(x_next_bit)));
    for (; it < val[0].begin() + circuit.size; ++it) it->clear();
}

void neuralNetwork::readConvWeight(i64 first_conv_id) {
    const i64 n = channel_out * channel_in * m * m;
    vector<double> dat(n);
    double mx = -10000, mn = 10000;
    for (auto &x : dat) {
        x = in->next();
        mx = max(mx, x);
        mn = min(mn, x);
    }
    w_bit = quantBits(mx, mn, Q);
    auto it = val[0].begin() + first_conv_id;
    for (double x : dat) *it++ = F((i64) (x * exp2(w_bit)));
}

void neuralNetwork::readBias(i64 first_bias_id) {
    auto it = val[0].begin() + first_bias_id;
    for (i64 co = 0; co < channel_out; ++co) *it++ = F((i64) (in->next() * exp2(w_bit + x_bit)));
}

void neuralNetwork::readFconWeight(i64 first_fc_id) {
    const i64 n = channel_out * channel_in;
    vector<double> dat(n);
    double mx = -10000, mn = 10000;
    for (auto &x : dat) {
        x = in->next();
        mx = max(mx, x);
        mn = min(mn, x);
    }
    w_bit = quantBits(mx, mn, Q);
    auto it = val[0].begin() + first_fc_id;
    for (double x : dat) *it++ = F((i64) (x * exp2(w_bit)));
}

// ---- auxiliary witnesses (:899-916) ------------------------------------------------------------------------------------------------------
// Every auxiliary input is a function of ONE earlier gate value; besides computing it, remember how (for zk_circuit_aux_ops): the layer
// being built is the source layer + 1, or, for the decomposition of a value that itself sits in val[0] (the window maxima), the step
// of the ops that produced it.  kind: 0 sign bit, 1 magnitude bit, 2 running maximum of max(0, value).
void neuralNetwork::recordAux(i64 src_layer, i64 src_idx, i64 dst_idx, u32 kind, i64 bit) {
    if (!aux_ops_) return;
    if (src_layer != 0) aux_step_ = src_layer + 1;
    if (aux_step_ <= 0 || aux_step_ >= (i64) aux_ops_->size() || bit < 0 || bit > 255) { aux_supported_ = false; return; }
    auto &ops = (*aux_ops_)[aux_step_];
    ops.push_back((uint32_t) src_idx);
    ops.push_back((uint32_t) dst_idx);
    ops.push_back((uint32_t) bit | (kind << 8) | (src_layer == 0 ? 1u << 10 : 0u));
}
void neuralNetwork::prepareDecmpBit(i64 layer_id, i64 idx, i64 dcmp_id, i64 bit_shift) {
    i64 data = std::abs(val[layer_id].at(idx).getInt64());
    val[0].at(dcmp_id) = F((i64) ((data >> bit_shift) & 1));
    recordAux(layer_id, idx, dcmp_id, 1, bit_shift);
}
void neuralNetwork::prepareFieldBit(const F &data, i64 dcmp_id, i64 bit_shift) {
    i64 tmp = std::abs(data.getInt64());
    val[0].at(dcmp_id) = F((i64) ((tmp >> bit_shift) & 1));
    aux_supported_ = false;   // (average pooling decomposes a window SUM: not one of the recorded op kinds; such models keep the host path)
}
void neuralNetwork::prepareSignBit(i64 layer_id, i64 idx, i64 dcmp_id) {
    val[0].at(dcmp_id) = val[layer_id].at(idx).isNegative() ? F_ONE : F_ZERO;
    recordAux(layer_id, idx, dcmp_id, 0, 0);
}
void neuralNetwork::prepareMax(i64 layer_id, i64 idx, i64 max_id) {
    F data = val[layer_id].at(idx).isNegative() ? F_ZERO : val[layer_id].at(idx);
    if (data > val[0].at(max_id)) val[0].at(max_id) = data;
    recordAux(layer_id, idx, max_id, 2, 0);
}

bool neuralNetwork::quantizeImage(const double *pixels, size_t n, vector<F> &out) const {
    const i64 n_pix = imagePixels();
    if ((i64) n < n_pix) throw std::invalid_argument("quantizeImage: too few pixel values");
    double mx = -10000, mn = 10000;
    for (i64 i = 0; i < n_pix; ++i) { mx = max(mx, pixels[i]); mn = min(mn, pixels[i]); }
    if (quantBits(mx, mn, Q) != image_bit_) return false;
    out.resize((size_t) n_pix * pic_parallel);
    auto it = out.begin();
    for (i64 p = 0; p < pic_parallel; ++p)
        for (i64 i = 0; i < n_pix; ++i) *it++ = F((i64) (pixels[i] * exp2(image_bit_)));
    return true;
}

bool neuralNetwork::scalesMatch(const uint64_t *ranges, size_t n_layers) const {
    for (const auto &d : scale_decisions_) {
        if ((size_t) d.layer >= n_layers) return false;
        const i64 x = (i64) (ranges[2 * d.layer] + ranges[2 * d.layer + 1]);   // getNextBit: (mx + mn).getInt64()
        const double real_scale = x / exp2(d.x_bit + d.w_bit);
        if ((int) log2(((1 << (Q - 1)) - 1) / real_scale) != d.next_bit) return false;
    }
    return true;
}

std::vector<uint8_t> neuralNetwork::scaleDecisionLayers(size_t n_layers) const {
    std::vector<uint8_t> f(n_layers, 0);
    for (const auto &d : scale_decisions_)
        if ((size_t) d.layer < n_layers) f[d.layer] = 1;
    return f;
}

void neuralNetwork::inferFromOutput(const vector<F> &output) {
    inferred.assign(pic_parallel, -1);
    if (full_conn.empty()) return;
    const int n_class = full_conn.back().channel_out;
    for (int p = 0; p < pic_parallel; ++p) {
        int k = -1;
        F best;
        for (int c = 0; c < n_class; ++c) {
            const F &tmp = output.at(matIdx(p, c, n_class));
            if (!tmp.isNegative() && (k == -1 || best < tmp)) { k = c; best = tmp; }
        }
        inferred[p] = k;
    }
}

// ---- circuit evaluation (:918-965) ----------------------------------------------------------------------------------------------------------
// out[g] += val[lu][u] * two_mul[sc]  /  val[Lu][u] * val[Lv][v] * two_mul[sc];  then the whole layer times `scale`.
// Field addition is exact, so the gates are spread over threads by output index without changing any value.
void neuralNetwork::calcNormalLayer(const layer &circuit, i64 layer_id) {
    vector<F> &out = val[layer_id];
    out.assign(circuit.size, F());
    const size_t n_gates = circuit.uni_gates.size() + circuit.bin_gates.size();
    unsigned nt = hostThreads > 0 ? (unsigned) hostThreads : std::max(1u, std::thread::hardware_concurrency());
    if (n_gates < (1u << 16)) nt = 1;
    nt = std::min(nt, 64u);
    const vector<F> &v0 = val[0];
    const vector<F> *prev = layer_id > 0 ? &val[layer_id - 1] : nullptr;
    auto work = [&](unsigned tid) {
        for (const auto &g : circuit.uni_gates) {
            if (nt > 1 && g.g % nt != tid) continue;
            const F &x = val[g.lu].at(g.u);
            out[g.g] += g.sc ? x * two_mul[g.sc] : x;
        }
        for (const auto &g : circuit.bin_gates) {
            if (nt > 1 && g.g % nt != tid) continue;
            const F &x = g.l ? prev->at(g.u) : v0.at(g.u);
            const F &y = (g.l & 1) ? (*prev)[g.v] : v0[g.v];
            F t = x * y;
            out[g.g] += g.sc ? t * two_mul[g.sc] : t;
        }
    };
    if (nt == 1) work(0);
    else {
        vector<std::thread> pool_;
        for (unsigned t = 0; t < nt; ++t) pool_.emplace_back(work, t);
        for (auto &t : pool_) t.join();
    }
    if (!circuit.scale.isOne())
        for (auto &x : out) x = x * circuit.scale;
}

// run body(first, last) over [0, n) on the host threads (blocks of a layer are independent)
static void parallelBlocks(size_t n, int hostThreads, const std::function<void(size_t, size_t)> &body) {
    unsigned nt = hostThreads > 0 ? (unsigned) hostThreads : std::max(1u, std::thread::hardware_concurrency());
    nt = (unsigned) std::min<size_t>(std::min(nt, 64u), std::max<size_t>(1, n));
    if (nt == 1) { body(0, n); return; }
    vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t] { body(n * t / nt, n * (t + 1) / nt); });
    for (auto &t : th) t.join();
}

// out[g block] += src[u block] .* src[v block] (:934-944); gates are grouped by output block so that threads never share one
void neuralNetwork::calcDotProdLayer(const layer &circuit, i64 layer_id) {
    vector<F> &out = val[layer_id];
    out.assign(circuit.size, F());
    const int fft_bit = circuit.fft_bit_length;
    const u32 fft_len = 1u << fft_bit;
    const vector<F> &src = val[layer_id - 1];
    const size_t n_out = circuit.size >> fft_bit;
    vector<u32> first(n_out + 1, 0), order(circuit.bin_gates.size());
    for (const auto &g : circuit.bin_gates) ++first.at(g.g + 1);
    for (size_t i = 0; i < n_out; ++i) first[i + 1] += first[i];
    {
        vector<u32> fill(first.begin(), first.end() - 1);
        for (u32 i = 0; i < circuit.bin_gates.size(); ++i) order[fill[circuit.bin_gates[i].g]++] = i;
    }
    parallelBlocks(n_out, hostThreads, [&](size_t b, size_t e) {
        for (size_t o = b; o < e; ++o)
            for (u32 k = first[o]; k < first[o + 1]; ++k) {
                const binGate &g = circuit.bin_gates[order[k]];
                const F *x = &src.at(((size_t) g.u << fft_bit) + fft_len - 1) - (fft_len - 1), *y = &src.at(((size_t) g.v << fft_bit) + fft_len - 1) - (fft_len - 1);
                F *d = &out[o << fft_bit];
                for (u32 s = 0; s < fft_len; ++s) d[s] += x[s] * y[s];
            }
    });
}

// radix-2 NTT of one block with shared bit-reversal and twiddle tables (the transform of fft() above, src/utils.cpp:105-145)
struct NttPlan {
    int logn;
    bool inverse;
    vector<u32> rev;
    vector<F> w;
    F ilen;
    NttPlan(int logn_, bool inv) : logn(logn_), inverse(inv), rev(1u << logn_), w(1u << logn_) {
        const u32 len = 1u << logn;
        rev[0] = 0;
        for (u32 i = 1; i < len; ++i) rev[i] = rev[i >> 1] >> 1 | (i & 1) << (logn - 1);
        w[0] = F_ONE;
        if (len > 1) {
            w[1] = getRootOfUnit(logn);
            if (inverse) F::inv(w[1], w[1]);
            for (u32 i = 2; i < len; ++i) w[i] = w[i - 1] * w[1];
        }
        F::inv(ilen, F((u64) len));
    }
    void run(F *arr) const {
        const u32 len = 1u << logn;
        for (u32 i = 0; i < len; ++i)
            if (rev[i] < i) std::swap(arr[i], arr[rev[i]]);
        for (u32 span = 2; span <= len; span <<= 1) {
            const u32 half = span >> 1, step = len / span;
            for (u32 j = 0; j < len; j += span)
                for (u32 k = 0; k < half; ++k) {
                    const F u = arr[j + k], v = arr[j + k + half] * w[step * k];
                    arr[j + k] = u + v;
                    arr[j + k + half] = u - v;
                }
        }
        if (inverse)
            for (u32 i = 0; i < len; ++i) arr[i] = arr[i] * ilen;
    }
};

void neuralNetwork::calcFFTLayer(const layer &circuit, i64 layer_id) {   // :946-965
    const i64 fft_len = 1LL << circuit.fft_bit_length, fft_lenh = fft_len >> 1;
    const bool inverse = circuit.ty == layerType::IFFT;
    vector<F> &out = val[layer_id];
    const vector<F> &src = val[layer_id - 1];
    out.assign(circuit.size, F());
    const NttPlan plan(circuit.fft_bit_length, inverse);
    const size_t n_blocks = inverse ? (circuit.size + fft_lenh - 1) / fft_lenh : (circuit.size + fft_len - 1) / fft_len;
    if ((inverse ? n_blocks * fft_len : n_blocks * fft_lenh) > src.size() || (inverse ? n_blocks * fft_lenh : n_blocks * fft_len) > out.size())
        throw std::out_of_range("calcFFTLayer: layer sizes are not whole FFT blocks");
    parallelBlocks(n_blocks, hostThreads, [&](size_t b, size_t e) {
        vector<F> arr(fft_len);
        for (size_t k = b; k < e; ++k) {
            if (!inverse) {   // half-length block, zero-extended, forward transform
                for (i64 j = 0; j < fft_lenh; ++j) arr[j] = src[k * fft_lenh + j];
                for (i64 j = fft_lenh; j < fft_len; ++j) arr[j].clear();
                plan.run(arr.data());
                for (i64 j = 0; j < fft_len; ++j) out[k * fft_len + j] = arr[j];
            } else {          // full block, inverse transform, keep the first half
                for (i64 j = 0; j < fft_len; ++j) arr[j] = src[k * fft_len + j];
                plan.run(arr.data());
                for (i64 j = 0; j < fft_lenh; ++j) out[k * fft_lenh + j] = arr[j];
            }
        }
    });
}

int neuralNetwork::getNextBit(int layer_id) {   // :967-977
    F mx = F_ZERO, mn = F_ZERO;
    for (const auto &x : val[layer_id]) {
        if (!x.isNegative()) mx = max(mx, x);
        else mn = max(mn, -x);
    }
    i64 x = (mx + mn).getInt64();
    double real_scale = x / exp2(x_bit + w_bit);
    return (int) log2(((1 << (Q - 1)) - 1) / real_scale);
}

void neuralNetwork::printInfer(prover &pr) {   // :994-1017
    inferred.assign(pic_parallel, -1);
    if (full_conn.empty()) return;
    const int n_class = full_conn.back().channel_out;
    for (int p = 0; p < pic_parallel; ++p) {
        int k = -1;
        F best;
        for (int c = 0; c < n_class; ++c) {
            const F &tmp = pr.val[SIZE - 1].at(matIdx(p, c, n_class));
            if (!tmp.isNegative() && (k == -1 || best < tmp)) {
                k = c;
                best = tmp;
            }
        }
        inferred[p] = k;
    }
    if (!o_file.empty()) {
        std::ofstream out(o_file);
        for (int k : inferred) out << k << std::endl;
    }
}

// ---- model zoo (src/models.cpp) -------------------------------------------------------------------------------------------------------------------
static convType pickConv(i64 kernel_size, i64 pparallel) { return kernel_size > 3 || pparallel > 1 ? FFT : NAIVE_FAST; }   // src/models.cpp:21

void vgg::configure(std::istream &config_in) {
    conv_section.resize(5);
    const i64 kernel_size = 3;
    i64 ch_in = pic_channel, new_nx = pic_size_x, new_ny = pic_size_y;
    const convType conv_ty = pickConv(kernel_size, pic_parallel);
    size_t idx = 0;
    string tok;
    while (config_in >> tok) {
        if (tok[0] != 'M' && tok[0] != 'A') {
            const i64 ch_out = std::stoi(tok);
            if (idx >= conv_section.size()) conv_section.resize(idx + 1);
            conv_section[idx].emplace_back(conv_ty, ch_out, ch_in, kernel_size);
            ch_in = ch_out;
        } else {
            ++idx;
            pool.emplace_back(tok[0] == 'M' ? MAX : AVG, 2, 1);
            new_nx = ((new_nx - pool.back().size) >> pool.back().stride_bl) + 1;
            new_ny = ((new_ny - pool.back().size) >> pool.back().stride_bl) + 1;
        }
    }
    full_conn.emplace_back(512, new_nx * new_ny * ch_in);
    full_conn.emplace_back(512, 512);
    full_conn.emplace_back(10, 512);
}

vgg::vgg(i64 psize_x, i64 psize_y, i64 pchannel, i64 pparallel, const string &i_filename, const string &c_filename,
         const std::string &o_filename, const std::string &n_filename)
    : neuralNetwork(psize_x, psize_y, pchannel, pparallel, i_filename, c_filename, o_filename) {
    if (n_filename.empty()) return;
    std::ifstream config_in(n_filename);
    if (!config_in.is_open()) throw std::runtime_error("vgg: cannot open the network description " + n_filename);
    configure(config_in);
}

std::unique_ptr<vgg> vgg::fromDescription(i64 psize, i64 pchannel, i64 pparallel, const std::string &description) {
    std::unique_ptr<vgg> nn(new vgg(psize, psize, pchannel, pparallel, "", "", "", ""));
    std::istringstream ss(description);
    nn->configure(ss);
    return nn;
}

static void vggFamily(vector<vector<convKernel>> &cs, vector<poolKernel> &pool, vector<fconKernel> &fc, const vector<int> &convs_per_section,
                      i64 pic_size, i64 pic_channel, i64 pparallel, poolType pool_ty) {
    const i64 start = 64, kernel_size = 3;
    const convType conv_ty = pickConv(kernel_size, pparallel);
    const i64 widths[5] = {start, start << 1, start << 2, start << 3, start << 3};
    cs.resize(5);
    i64 ch_in = pic_channel, nx = pic_size;
    for (int s = 0; s < 5; ++s) {
        for (int k = 0; k < convs_per_section[s]; ++k) {
            cs[s].emplace_back(conv_ty, widths[s], ch_in, kernel_size);
            ch_in = widths[s];
        }
        pool.emplace_back(pool_ty, 2, 1);
        nx = ((nx - 2) >> 1) + 1;
    }
    if (pic_size == 224) {
        fc.emplace_back(4096, nx * nx * (start << 3));
        fc.emplace_back(4096, 4096);
        fc.emplace_back(1000, 4096);
    } else {
        fc.emplace_back(512, nx * nx * (start << 3));
        fc.emplace_back(512, 512);
        fc.emplace_back(10, 512);
    }
}

vgg16::vgg16(i64 psize_x, i64 psize_y, i64 pchannel, i64 pparallel, poolType pool_ty_, const std::string &i_filename, const string &c_filename,
             const std::string &o_filename)
    : neuralNetwork(psize_x, psize_y, pchannel, pparallel, i_filename, c_filename, o_filename) {
    vggFamily(conv_section, pool, full_conn, {2, 2, 3, 3, 3}, pic_size_x, pic_channel, pparallel, pool_ty_);   // src/models.cpp:43-96
}

vgg11::vgg11(i64 psize_x, i64 psize_y, i64 pchannel, i64 pparallel, poolType pool_ty_, const std::string &i_filename, const string &c_filename,
             const std::string &o_filename)
    : neuralNetwork(psize_x, psize_y, pchannel, pparallel, i_filename, c_filename, o_filename) {
    vggFamily(conv_section, pool, full_conn, {1, 1, 2, 2, 2}, pic_size_x, pic_channel, pparallel, pool_ty_);   // src/models.cpp:98-146
}

lenet::lenet(i64 psize_x, i64 psize_y, i64 pchannel, i64 pparallel, poolType pool_ty_, const std::string &i_filename, const string &c_filename,
             const std::string &o_filename)
    : neuralNetwork(psize_x, psize_y, pchannel, pparallel, i_filename, c_filename, o_filename) {   // src/models.cpp:166-186
    const i64 kernel_size = 5;
    const convType conv_ty = pickConv(kernel_size, pparallel);
    conv_section.resize(2);
    conv_section[0].emplace_back(conv_ty, 6, pchannel, kernel_size, 0, psize_x == 28 && psize_y == 28 ? 2 : 0);
    pool.emplace_back(pool_ty_, 2, 1);
    conv_section[1].emplace_back(conv_ty, 16, 6, kernel_size, 0, 0);
    pool.emplace_back(pool_ty_, 2, 1);
    full_conn.emplace_back(120, 400);
    full_conn.emplace_back(84, 120);
    full_conn.emplace_back(10, 84);
}

lenetCifar::lenetCifar(i64 psize_x, i64 psize_y, i64 pchannel, i64 pparallel, poolType pool_ty_, const std::string &i_filename,
                       const string &c_filename, const std::string &o_filename)
    : neuralNetwork(psize_x, psize_y, pchannel, pparallel, i_filename, c_filename, o_filename) {   // src/models.cpp:188-206
    const i64 kernel_size = 5;
    const convType conv_ty = pickConv(kernel_size, pparallel);
    conv_section.resize(3);
    conv_section[0].emplace_back(conv_ty, 6, pchannel, kernel_size, 0, 0);
    pool.emplace_back(pool_ty_, 2, 1);
    conv_section[1].emplace_back(conv_ty, 16, 6, kernel_size, 0, 0);
    pool.emplace_back(pool_ty_, 2, 1);
    conv_section[2].emplace_back(conv_ty, 120, 16, kernel_size, 0, 0);
    full_conn.emplace_back(84, 120);
    full_conn.emplace_back(10, 84);
}

ccnn::ccnn(i64 psize_x, i64 psize_y, i64 pparallel, i64 pchannel, poolType pool_ty_)
    : neuralNetwork(psize_x, psize_y, pchannel, pparallel, "", "", "") {   // src/models.cpp:148-164
    const i64 kernel_size = 2;
    conv_section.resize(1);
    conv_section[0].emplace_back(pickConv(kernel_size, pparallel), 2, pchannel, kernel_size, 0, 0);
    pool.emplace_back(pool_ty_, 2, 1);
}

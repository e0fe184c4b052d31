/* oracle/opv_oracle.c — TEST INFRASTRUCTURE ONLY (see opv_oracle.h).
 *
 * Plain-C restatement of the reference receive chain, operation by operation, so
 * that its doubles are bit-identical to the reference's (same libm, same order of
 * IEEE operations; build with -ffp-contract=off).  Citations: file:line in
 * /root/reference/src/opv-demod.cpp ("mod:" = opv-mod.cpp).
 */
#include "opv_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SAMPLE_RATE 2168000.0              /* :40 */
#define SYMBOL_RATE (SAMPLE_RATE / 40.0)   /* :41 */
#define FREQ_DEV 13550.0                   /* :42 */
#define PI 3.14159265358979323846          /* :43 */
#define TWO_PI (2.0 * PI)                  /* :44 */
#define SYNC_WORD 0x02B8DBu                /* :46 */
#define SOFT_MAX 7                         /* :54 */
#define G1_MASK 0x4F                       /* :55 */
#define G2_MASK 0x6D                       /* :56 */
#define NUM_STATES 64                      /* :57 */
#define SYNC_MISS_LIMIT 5                  /* :60 */

static double clampd(double v, double lo, double hi) { /* std::clamp */
    return (v < lo) ? lo : (hi < v) ? hi : v;
}

/* ------------------------------------------------------------------ A2 */
/* one candidate of estimate_offset (:136-159 == :170-193) */
static double candidate_energy(const int16_t* iq, size_t n, double offset) {
    double phase_f1 = 0, phase_f2 = 0;
    double phase_inc_f1 = TWO_PI * (-FREQ_DEV + offset) / SAMPLE_RATE;
    double phase_inc_f2 = TWO_PI * (+FREQ_DEV + offset) / SAMPLE_RATE;
    double total_energy = 0;
    size_t test_samples = n < (size_t)(ORA_SPS * 1000) ? n : (size_t)(ORA_SPS * 1000);
    for (size_t sym = 0; sym < test_samples / ORA_SPS; ++sym) {
        double c1r = 0, c1i = 0, c2r = 0, c2i = 0;
        for (size_t i = 0; i < ORA_SPS; ++i) {
            size_t idx = sym * ORA_SPS + i;
            double a = iq[2 * idx], b = iq[2 * idx + 1];
            double l1c = cos(phase_f1), l1s = sin(phase_f1);
            double l2c = cos(phase_f2), l2s = sin(phase_f2);
            /* s * conj(lo) = (a*c + b*s, b*c - a*s)  (complex product, no contraction) */
            c1r += a * l1c - b * (-l1s);
            c1i += a * (-l1s) + b * l1c;
            c2r += a * l2c - b * (-l2s);
            c2i += a * (-l2s) + b * l2c;
            phase_f1 += phase_inc_f1;
            phase_f2 += phase_inc_f2;
        }
        total_energy += (c1r * c1r + c1i * c1i) + (c2r * c2r + c2i * c2i);
    }
    return total_energy;
}

double ora_estimate_offset(const int16_t* iq, size_t n) { /* :131-202 */
    double best_offset = 0, best_energy = 0;
    for (double offset = -1500; offset <= 1500; offset += 25) {
        double e = candidate_energy(iq, n, offset);
        if (e > best_energy) { best_energy = e; best_offset = offset; }
    }
    double fine_best = best_offset;
    for (double offset = best_offset - 30; offset <= best_offset + 30; offset += 5) {
        double e = candidate_energy(iq, n, offset);
        if (e > best_energy) { best_energy = e; fine_best = offset; }
    }
    return fine_best;
}

/* ------------------------------------------------------------------ coherent (-c) */
void ora_codemod_init(ora_codemod_t* d) { /* :367-377 */
    memset(d, 0, sizeof(*d));
    d->afc_alpha = 0.001; d->pll_alpha = 0.01; d->pll_beta = 0.001;
}

void ora_codemod_set_pll_bandwidth(ora_codemod_t* d, double bw) { /* :561-568 */
    double wn = bw * TWO_PI;
    double zeta = 0.707;
    d->pll_alpha = 2.0 * zeta * wn / SYMBOL_RATE;
    d->pll_beta = wn * wn / (SYMBOL_RATE * SYMBOL_RATE);
}

/* TEST HOOK (not in the reference): a perturbation of the Costas loop's initial phase, in radians.  The loop is
 * chaotic on these signals (it never locks), so a last-bit difference in sin/cos grows until the trajectory is a
 * different one; tests use this to measure that horizon on the oracle itself and to tell which captures have
 * frames that are stable under it (tests/test_oracle.py::test_coherent_sensitivity). */
static double g_codemod_perturb = 0.0;
void ora_set_codemod_perturb(double rad) { g_codemod_perturb = rad; }

size_t ora_codemod_demodulate(ora_codemod_t* d, const int16_t* iq, size_t n, double* soft_out, size_t cap) { /* :450-548 */
    size_t ns = 0;
    d->carrier_phase += g_codemod_perturb;
    double phase_inc_f1 = TWO_PI * (-FREQ_DEV + d->freq_offset) / SAMPLE_RATE;
    double phase_inc_f2 = TWO_PI * (+FREQ_DEV + d->freq_offset) / SAMPLE_RATE;
    for (size_t sym = 0; sym < n / ORA_SPS; ++sym) {
        double c1r = 0, c1i = 0, c2r = 0, c2i = 0;
        for (size_t i = 0; i < ORA_SPS; ++i) { /* :461-483 */
            size_t idx = sym * ORA_SPS + i;
            double a = iq[2 * idx], b = iq[2 * idx + 1];
            double pr = cos(d->carrier_phase), pi = -sin(d->carrier_phase);
            double xr = a * pr - b * pi, xi = a * pi + b * pr;          /* samples[idx] * phase_rot */
            double l1c = cos(d->phase_f1), l1s = sin(d->phase_f1);
            double l2c = cos(d->phase_f2), l2s = sin(d->phase_f2);
            c1r += xr * l1c - xi * (-l1s);                               /* corrected * conj(lo) */
            c1i += xr * (-l1s) + xi * l1c;
            c2r += xr * l2c - xi * (-l2s);
            c2i += xr * (-l2s) + xi * l2c;
            d->phase_f1 += phase_inc_f1;
            d->phase_f2 += phase_inc_f2;
            d->carrier_phase += d->loop_freq;
        }
        while (d->phase_f1 > PI) d->phase_f1 -= TWO_PI;                  /* :486-491 */
        while (d->phase_f1 < -PI) d->phase_f1 += TWO_PI;
        while (d->phase_f2 > PI) d->phase_f2 -= TWO_PI;
        while (d->phase_f2 < -PI) d->phase_f2 += TWO_PI;
        while (d->carrier_phase > PI) d->carrier_phase -= TWO_PI;
        while (d->carrier_phase < -PI) d->carrier_phase += TWO_PI;
        double e1 = c1r * c1r + c1i * c1i, e2 = c2r * c2r + c2i * c2i;   /* :494-495 */
        if (ns < cap) soft_out[ns] = c2r - c1r;                          /* :500-505 */
        ++ns;
        double dr = (e1 > e2) ? c1r : c2r, di = (e1 > e2) ? c1i : c2i;   /* :510 */
        double mag = hypot(dr, di);                                      /* std::abs */
        double phase_error = 0;
        if (mag > 1e-10) phase_error = di / mag;                         /* :514-520 */
        d->loop_freq += d->pll_beta * phase_error;                       /* :524-525 */
        d->carrier_phase += d->pll_alpha * phase_error;
        d->loop_freq = clampd(d->loop_freq, -0.1, 0.1);                  /* :528 */
        if (sym > 0) {                                                   /* :533-541 */
            double npi = -d->prev_im;                                    /* dominant * conj(prev_dominant) */
            double xr = dr * d->prev_re - di * npi;
            double xi = dr * npi + di * d->prev_re;
            double phase_diff = atan2(xi, xr);
            double freq_err = phase_diff * SYMBOL_RATE / TWO_PI;
            d->freq_offset += d->afc_alpha * freq_err;
            d->freq_offset = clampd(d->freq_offset, -2000.0, 2000.0);
            phase_inc_f1 = TWO_PI * (-FREQ_DEV + d->freq_offset) / SAMPLE_RATE;
            phase_inc_f2 = TWO_PI * (+FREQ_DEV + d->freq_offset) / SAMPLE_RATE;
        }
        d->prev_re = dr; d->prev_im = di;                                /* :543 */
    }
    return ns;
}

/* ------------------------------------------------------------------ A3 */
void ora_demod_init(ora_demod_t* d) { /* :110-119 */
    memset(d, 0, sizeof(*d));
    d->afc_alpha = 0.001;
    d->alpha_timing = 0.005;
    d->beta_timing = 0.00001;
}

static void interp(const int16_t* iq, double idx, size_t len, double* re, double* im) { /* :122-128 */
    if (idx < 0) idx = 0;
    if (idx >= (double)(len - 1)) idx = (double)(len - 2);
    size_t i = (size_t)idx;
    double f = idx - (double)i;
    double w0 = 1.0 - f;
    *re = (double)iq[2 * i] * w0 + (double)iq[2 * i + 2] * f;
    *im = (double)iq[2 * i + 1] * w0 + (double)iq[2 * i + 3] * f;
}

size_t ora_demodulate(ora_demod_t* d, const int16_t* iq, size_t n, double* soft_out, size_t cap) { /* :206-329 */
    size_t n_soft = 0;
    double phase_inc_f1 = TWO_PI * (-FREQ_DEV + d->freq_offset) / SAMPLE_RATE;
    double phase_inc_f2 = TWO_PI * (+FREQ_DEV + d->freq_offset) / SAMPLE_RATE;
    const double EL_OFFSET = ORA_SPS / 4.0;
    double pos = d->mu;

    while (pos + (double)ORA_SPS + EL_OFFSET < (double)n) {
        double c1r = 0, c1i = 0, c2r = 0, c2i = 0;
        double e1r = 0, e1i = 0, e2r = 0, e2i = 0;
        double l1r = 0, l1i = 0, l2r = 0, l2i = 0;
        double ph1 = d->phase_f1, ph2 = d->phase_f2;

        for (size_t i = 0; i < ORA_SPS; ++i) {
            double p_on = pos + (double)i;
            double p_early = p_on - EL_OFFSET;
            double p_late = p_on + EL_OFFSET;
            double onr, oni, er, ei, lr, li;
            interp(iq, p_on, n, &onr, &oni);
            if (p_early >= 0) interp(iq, p_early, n, &er, &ei);
            else { er = iq[0]; ei = iq[1]; }
            interp(iq, p_late, n, &lr, &li);

            double c1 = cos(ph1), s1 = sin(ph1);
            double c2 = cos(ph2), s2 = sin(ph2);

            c1r += onr * c1 - oni * (-s1);  c1i += onr * (-s1) + oni * c1;
            c2r += onr * c2 - oni * (-s2);  c2i += onr * (-s2) + oni * c2;
            e1r += er * c1 - ei * (-s1);    e1i += er * (-s1) + ei * c1;
            e2r += er * c2 - ei * (-s2);    e2i += er * (-s2) + ei * c2;
            l1r += lr * c1 - li * (-s1);    l1i += lr * (-s1) + li * c1;
            l2r += lr * c2 - li * (-s2);    l2i += lr * (-s2) + li * c2;

            ph1 += phase_inc_f1;
            ph2 += phase_inc_f2;
        }
        d->phase_f1 = ph1;
        d->phase_f2 = ph2;
        while (d->phase_f1 > PI) d->phase_f1 -= TWO_PI;
        while (d->phase_f1 < -PI) d->phase_f1 += TWO_PI;
        while (d->phase_f2 > PI) d->phase_f2 -= TWO_PI;
        while (d->phase_f2 < -PI) d->phase_f2 += TWO_PI;

        double e1 = c1r * c1r + c1i * c1i;
        double e2 = c2r * c2r + c2i * c2i;
        if (n_soft < cap && soft_out) soft_out[n_soft] = e2 - e1;
        ++n_soft;

        double ted;
        if (e1 > e2) {
            double ee = e1r * e1r + e1i * e1i;
            double el = l1r * l1r + l1i * l1i;
            ted = (el - ee) / (el + ee + 1e-10);
        } else {
            double ee = e2r * e2r + e2i * e2i;
            double el = l2r * l2r + l2i * l2i;
            ted = (el - ee) / (el + ee + 1e-10);
        }
        d->timing_freq += d->beta_timing * ted;
        d->timing_freq = clampd(d->timing_freq, -0.1, 0.1);
        double timing_adj = d->alpha_timing * ted + d->timing_freq;
        timing_adj = clampd(timing_adj, -2.0, 2.0);

        if (n_soft > 1) { /* :289 — skipped for the first symbol of every call */
            double dr, di, pr, pi_;
            if (e1 > e2) { dr = c1r; di = c1i; pr = d->prev1_re; pi_ = d->prev1_im; }
            else         { dr = c2r; di = c2i; pr = d->prev2_re; pi_ = d->prev2_im; }
            /* dom * conj(prev) */
            double xr = dr * pr - di * (-pi_);
            double xi = dr * (-pi_) + di * pr;
            double pd = atan2(xi, xr);
            double ferr = pd * SYMBOL_RATE / TWO_PI;
            d->freq_offset += d->afc_alpha * ferr;
            d->freq_offset = clampd(d->freq_offset, -2000.0, 2000.0);
            phase_inc_f1 = TWO_PI * (-FREQ_DEV + d->freq_offset) / SAMPLE_RATE;
            phase_inc_f2 = TWO_PI * (+FREQ_DEV + d->freq_offset) / SAMPLE_RATE;
        }
        d->prev1_re = c1r; d->prev1_im = c1i;
        d->prev2_re = c2r; d->prev2_im = c2i;

        pos += (double)ORA_SPS + timing_adj;
    }
    size_t used = (size_t)pos;
    d->mu = pos - (double)used;
    d->leftover = n - used;
    return n_soft;
}

/* ------------------------------------------------------------------ A5 */
static double sync_pattern(int i) { /* :597-600 */
    int bit = (SYNC_WORD >> (ORA_SYNC_BITS - 1 - i)) & 1;
    return bit ? -1.0 : +1.0;
}

void ora_tracker_init(ora_tracker_t* t) { memset(t, 0, sizeof(*t)); t->state = ORA_HUNTING; }

static double soft_correlate(const ora_tracker_t* t, double* raw) { /* :743-757 */
    double sum = 0.0, energy = 0.0;
    for (size_t i = 0; i < ORA_SYNC_BITS; ++i) {
        size_t b = (t->corr_idx + i) % ORA_SYNC_BITS;
        double s = t->corr_buf[b];
        sum += s * sync_pattern((int)i);
        energy += fabs(s);
    }
    if (raw) *raw = sum;
    if (energy < 100.0) return 0.0;
    return sum / energy;
}

static void push_ev(ora_event_t* ev, int* n_ev, int type, int count, size_t idx, double corr, double raw) {
    if (!ev || !n_ev) return;
    ora_event_t* e = &ev[(*n_ev)++];
    e->type = type; e->count = count; e->sym_idx = (int64_t)idx; e->corr = corr; e->raw = raw;
}

int ora_tracker_process(ora_tracker_t* t, double soft, size_t sym_idx, double* payload_out,
                        double* quality_out, ora_event_t* ev, int* n_ev) { /* :615-736 */
    int ready = 0;
    if (n_ev) *n_ev = 0;
    t->corr_buf[t->corr_idx] = soft;
    t->corr_idx = (t->corr_idx + 1) % ORA_SYNC_BITS;
    t->total_symbols++;
    if (t->collecting && t->n_pending < ORA_ENCODED_BITS) t->pending[t->n_pending++] = soft;
    t->since_sync++;

    switch (t->state) {
    case ORA_HUNTING: {
        if (t->total_symbols < ORA_SYNC_BITS) break;
        double raw, norm = soft_correlate(t, &raw);
        if (raw >= 5000.0 && norm >= 0.85) {
            t->state = ORA_VERIFYING;
            t->sync_quality = norm;
            t->since_sync = 0;
            t->collecting = 1;
            t->n_pending = 0;
            push_ev(ev, n_ev, ORA_EV_HUNT_TO_VERIFY, 0, sym_idx, norm, raw);
        }
        break;
    }
    case ORA_VERIFYING: {
        if (t->since_sync >= ORA_ENCODED_BITS) {
            ready = 1;
            if (quality_out) *quality_out = t->sync_quality;
            if (payload_out) memcpy(payload_out, t->pending, sizeof(double) * ORA_ENCODED_BITS);
            t->total_frames++;
            t->n_pending = 0;
            t->collecting = 0;
            t->state = ORA_LOCKED;
            t->misses = 0;
            push_ev(ev, n_ev, ORA_EV_VERIFY_TO_LOCKED, t->total_frames, sym_idx, 0, 0);
        }
        break;
    }
    case ORA_LOCKED: {
        if (t->since_sync == ORA_FRAME_SYMBOLS) {
            double raw, corr = soft_correlate(t, &raw);
            if (corr >= 0.70) {
                t->misses = 0;
                t->sync_quality = corr;
                t->collecting = 1;
                t->n_pending = 0;
                push_ev(ev, n_ev, ORA_EV_SYNC_OK, 0, sym_idx, corr, raw);
            } else {
                t->misses++;
                push_ev(ev, n_ev, ORA_EV_SYNC_MISS, t->misses, sym_idx, corr, raw);
                if (t->misses >= SYNC_MISS_LIMIT) {
                    t->state = ORA_HUNTING;
                    t->collecting = 0;
                    push_ev(ev, n_ev, ORA_EV_LOST_LOCK, 0, sym_idx, 0, 0);
                    break;
                }
                t->sync_quality = corr;
                t->collecting = 1;
                t->n_pending = 0;
            }
            t->since_sync = 0;
        }
        if (t->collecting && t->n_pending >= ORA_ENCODED_BITS) {
            ready = 1;
            if (quality_out) *quality_out = t->sync_quality;
            if (payload_out) memcpy(payload_out, t->pending, sizeof(double) * ORA_ENCODED_BITS);
            t->total_frames++;
            t->n_pending = 0;
            t->collecting = 0;
        }
        break;
    }
    }
    return ready;
}

/* ------------------------------------------------------------------ A6-A8 */
size_t ora_deinterleave_addr(size_t idx) { /* :792-795 */
    size_t pos = (idx % 32) * 67 + (idx / 32);
    return (pos / 8) * 8 + (7 - pos % 8);
}

int ora_viterbi_decode(const int32_t* soft_in, uint8_t* bits_out) { /* :802-846 */
    int metrics[NUM_STATES], next[NUM_STATES];
    uint8_t (*decisions)[NUM_STATES] = malloc((size_t)ORA_FRAME_BITS * NUM_STATES);
    for (int s = 0; s < NUM_STATES; ++s) metrics[s] = 0x7FFFFFFF;
    metrics[0] = 0;
    for (size_t t = 0; t < ORA_FRAME_BITS; ++t) {
        int sg1 = soft_in[t * 2], sg2 = soft_in[t * 2 + 1];
        for (int s = 0; s < NUM_STATES; ++s) {
            int p0 = s / 2, p1 = p0 + 32;
            int in = s % 2;
            int f0 = (in << 6) | p0, f1 = (in << 6) | p1;
            int e1_0 = __builtin_parity(f0 & G1_MASK), e2_0 = __builtin_parity(f0 & G2_MASK);
            int e1_1 = __builtin_parity(f1 & G1_MASK), e2_1 = __builtin_parity(f1 & G2_MASK);
            int bm0 = (e1_0 ? SOFT_MAX - sg1 : sg1) + (e2_0 ? SOFT_MAX - sg2 : sg2);
            int bm1 = (e1_1 ? SOFT_MAX - sg1 : sg1) + (e2_1 ? SOFT_MAX - sg2 : sg2);
            int m0 = (metrics[p0] < 0x7FFFFFF0) ? metrics[p0] + bm0 : 0x7FFFFFFF;
            int m1 = (metrics[p1] < 0x7FFFFFF0) ? metrics[p1] + bm1 : 0x7FFFFFFF;
            if (m0 <= m1) { next[s] = m0; decisions[t][s] = 0; }
            else          { next[s] = m1; decisions[t][s] = 1; }
        }
        memcpy(metrics, next, sizeof(metrics));
    }
    int best = 0;
    for (int s = 1; s < NUM_STATES; ++s)
        if (metrics[s] < metrics[best]) best = s;
    int s = best;
    for (int t = ORA_FRAME_BITS - 1; t >= 0; --t) {
        bits_out[t] = (uint8_t)(s % 2);
        s = (decisions[t][s] == 0) ? s / 2 : s / 2 + 32;
    }
    free(decisions);
    return metrics[best];
}

void ora_lfsr_table(uint8_t* out) { /* :887-893 */
    uint8_t lfsr = 0xFF;
    for (size_t i = 0; i < ORA_FRAME_BYTES; ++i) {
        uint8_t r = 0;
        for (int b = 7; b >= 0; --b) {
            r |= (uint8_t)(((lfsr >> 7) & 1) << b);
            lfsr = (uint8_t)((lfsr << 1) | (((lfsr >> 7) ^ (lfsr >> 6) ^ (lfsr >> 4) ^ (lfsr >> 2)) & 1));
        }
        out[i] = r;
    }
}

void ora_quantise(const double* soft, int32_t* q, double* scale_out) { /* :856-866 */
    double scale = 0;
    for (size_t i = 0; i < ORA_ENCODED_BITS; ++i) scale += fabs(soft[i]);
    scale /= ORA_ENCODED_BITS;
    if (scale_out) *scale_out = scale;
    if (scale < 1e-10) { for (size_t i = 0; i < ORA_ENCODED_BITS; ++i) q[i] = -1; return; }
    for (size_t i = 0; i < ORA_ENCODED_BITS; ++i) {
        double n = (-soft[i] / scale) * 3.5 + 3.5;
        int v = (int)(n + 0.5);
        q[i] = v < 0 ? 0 : (v > SOFT_MAX ? SOFT_MAX : v);
    }
}

int ora_frame_decode(const double* soft, uint8_t* out) { /* :854-898 */
    int32_t qs[ORA_ENCODED_BITS], deint[ORA_ENCODED_BITS];
    double scale;
    ora_quantise(soft, qs, &scale);
    if (scale < 1e-10) return -1;
    for (size_t i = 0; i < ORA_ENCODED_BITS; ++i) deint[i] = qs[ora_deinterleave_addr(i)];
    uint8_t bits[ORA_FRAME_BITS];
    int metric = ora_viterbi_decode(deint, bits);
    uint8_t lfsr[ORA_FRAME_BYTES];
    ora_lfsr_table(lfsr);
    for (size_t i = 0; i < ORA_FRAME_BYTES; ++i) {
        uint8_t b = 0;
        for (int j = 0; j < 8; ++j) b |= (uint8_t)(bits[ORA_FRAME_BITS - 1 - i * 8 - j] << j);
        out[i] = b ^ lfsr[i];
    }
    return metric;
}

/* ------------------------------------------------------------------ main() drivers */
typedef struct {
    ora_tracker_t tracker;
    size_t total_symbols;
    ora_result_t* res;
} run_ctx_t;

static void feed_soft(run_ctx_t* c, const double* soft, size_t n) { /* :1045-1065 == :1186-1205 */
    ora_result_t* r = c->res;
    double payload[ORA_ENCODED_BITS]; /* on the stack: the tests run many captures on parallel threads */
    for (size_t i = 0; i < n; ++i) {
        ora_event_t ev[2]; int n_ev = 0; double q;
        int ready = ora_tracker_process(&c->tracker, soft[i], c->total_symbols + i, payload, &q, ev, &n_ev);
        for (int k = 0; k < n_ev; ++k) {
            if (r->events && r->n_events < r->cap_events) r->events[r->n_events] = ev[k];
            r->n_events++;
        }
        if (ready) {
            uint8_t frame[ORA_FRAME_BYTES];
            int metric = ora_frame_decode(payload, frame);
            if (metric >= 0) {
                if (r->n_frames < r->cap_frames) {
                    if (r->frames) memcpy(r->frames + r->n_frames * ORA_FRAME_BYTES, frame, ORA_FRAME_BYTES);
                    if (r->metrics) r->metrics[r->n_frames] = metric;
                    if (r->frame_ready_idx) r->frame_ready_idx[r->n_frames] = (int64_t)(c->total_symbols + i);
                }
                r->n_frames++;
                if (metric == 0) r->n_perfect++;
            } else {
                r->n_dropped++;
            }
        }
    }
    c->total_symbols += n;
}

int ora_run(const ora_cfg_t* cfg, const int16_t* iq, size_t n, ora_result_t* res) {
    res->n_frames = res->n_perfect = res->n_dropped = 0;
    res->n_soft = res->n_events = res->n_chunks = 0;
    res->est_offset = 0; res->total_samples = 0;
    run_ctx_t ctx; memset(&ctx, 0, sizeof(ctx));
    ora_tracker_init(&ctx.tracker);
    ctx.res = res;
    ora_demod_t d; ora_demod_init(&d);
    size_t cap_tmp = n / ORA_SPS + 16;
    double* soft = malloc(sizeof(double) * (cap_tmp ? cap_tmp : 1));
    if (!soft) return -1;

    if (!cfg->streaming && cfg->coherent) { /* :1144-1161 */
        ora_codemod_t cd; ora_codemod_init(&cd);
        res->est_offset = ora_estimate_offset(iq, n);   /* CoherentMSKDemodulator::estimate_offset (:379-446) is the same search */
        cd.freq_offset = res->est_offset;
        cd.afc_alpha = cfg->afc_alpha;
        ora_codemod_set_pll_bandwidth(&cd, cfg->pll_bw);
        if (res->chunk_starts && res->cap_chunks) res->chunk_starts[0] = 0;
        res->n_chunks = 1;
        size_t ns = ora_codemod_demodulate(&cd, iq, n, soft, cap_tmp);
        for (size_t i = 0; i < ns; ++i) { if (res->soft && res->n_soft < res->cap_soft) res->soft[res->n_soft] = soft[i]; res->n_soft++; }
        feed_soft(&ctx, soft, ns);
        d.freq_offset = cd.freq_offset;
    } else if (!cfg->streaming) { /* :1127-1216 */
        res->est_offset = ora_estimate_offset(iq, n);
        d.freq_offset = res->est_offset;
        d.afc_alpha = cfg->afc_alpha;
        if (res->chunk_starts && res->cap_chunks) res->chunk_starts[0] = 0;
        res->n_chunks = 1;
        size_t ns = ora_demodulate(&d, iq, n, soft, cap_tmp);
        for (size_t i = 0; i < ns; ++i) { if (res->soft && res->n_soft < res->cap_soft) res->soft[res->n_soft] = soft[i]; res->n_soft++; }
        feed_soft(&ctx, soft, ns);
    } else { /* :995-1125 */
        if (cfg->have_init_offset) d.freq_offset = cfg->init_offset;
        d.afc_alpha = cfg->afc_alpha;
        size_t start = 0;      /* global index of chunk_buf[0] */
        int first = 1;
        /* chunk_buf always reaches exactly CHUNK samples (one push at a time, :1022-1026) */
        while (n - start >= ORA_CHUNK_SAMPLES) {
            const int16_t* c = iq + 2 * start;
            res->total_samples += ORA_CHUNK_SAMPLES;
            if (first) {
                if (!cfg->have_init_offset) {
                    res->est_offset = ora_estimate_offset(c, ORA_CHUNK_SAMPLES);
                    d.freq_offset = res->est_offset;
                }
                first = 0;
            }
            if (res->chunk_starts && res->n_chunks < res->cap_chunks) res->chunk_starts[res->n_chunks] = (int64_t)start;
            res->n_chunks++;
            size_t ns = ora_demodulate(&d, c, ORA_CHUNK_SAMPLES, soft, cap_tmp);
            for (size_t i = 0; i < ns; ++i) { if (res->soft && res->n_soft < res->cap_soft) res->soft[res->n_soft] = soft[i]; res->n_soft++; }
            feed_soft(&ctx, soft, ns);
            size_t leftover = d.leftover;
            if (leftover > 0 && leftover < ORA_CHUNK_SAMPLES) start += ORA_CHUNK_SAMPLES - leftover; /* :1071-1073 */
            else start += ORA_CHUNK_SAMPLES;                                                       /* :1075 */
        }
        if (n - start > 0) { /* :1088-1113 */
            if (res->chunk_starts && res->n_chunks < res->cap_chunks) res->chunk_starts[res->n_chunks] = (int64_t)start;
            res->n_chunks++;
            size_t ns = ora_demodulate(&d, iq + 2 * start, n - start, soft, cap_tmp);
            for (size_t i = 0; i < ns; ++i) { if (res->soft && res->n_soft < res->cap_soft) res->soft[res->n_soft] = soft[i]; res->n_soft++; }
            feed_soft(&ctx, soft, ns);
        }
    }
    res->final_freq = d.freq_offset;
    res->final_tfreq = d.timing_freq;
    res->final_state = ctx.tracker.state;
    free(soft);
    return 0;
}

/* ------------------------------------------------------------------ TX chain (capture generator) */
static int b40_digit(char c) { /* mod:82-90 */
    if (c >= 'A' && c <= 'Z') return c - 'A' + 1;
    if (c >= 'a' && c <= 'z') return c - 'a' + 1;
    if (c >= '0' && c <= '9') return c - '0' + 27;
    if (c == '-') return 37;
    if (c == '/') return 38;
    if (c == '.') return 39;
    return 0;
}

void ora_base40_encode(const char* cs, uint8_t* out) { /* mod:63-79 */
    uint64_t v = 0;
    for (int i = (int)strlen(cs) - 1; i >= 0; --i) { v *= 40; v += (uint64_t)b40_digit(cs[i]); }
    for (int k = 0; k < 6; ++k) out[k] = (uint8_t)((v >> (40 - 8 * k)) & 0xFF);
}

void ora_bert_frame(const char* cs, uint32_t token, uint32_t frame_num, uint8_t* f) { /* mod:339-361 */
    memset(f, 0, ORA_FRAME_BYTES);
    ora_base40_encode(cs, f);
    f[6] = (token >> 16) & 0xFF; f[7] = (token >> 8) & 0xFF; f[8] = token & 0xFF;
    for (size_t i = 0; i < ORA_FRAME_BYTES - 12; ++i) f[12 + i] = (uint8_t)((frame_num + i) & 0xFF);
}

void ora_encode_frame(const uint8_t* payload, uint8_t* out_bits) { /* mod:159-213 */
    uint8_t lfsr[ORA_FRAME_BYTES], enc[ORA_ENCODED_BITS];
    ora_lfsr_table(lfsr); /* same generator as mod:97-113 */
    uint8_t sr = 0;
    size_t o = 0;
    for (int byte_idx = ORA_FRAME_BYTES - 1; byte_idx >= 0; --byte_idx) {
        uint8_t byte = payload[byte_idx] ^ lfsr[byte_idx];
        for (int bit_pos = 7; bit_pos >= 0; --bit_pos) {
            uint8_t in = (byte >> bit_pos) & 1;
            uint8_t st = (uint8_t)((in << 6) | sr);              /* mod:125 */
            enc[o++] = (uint8_t)__builtin_parity(st & 0x4F);
            enc[o++] = (uint8_t)__builtin_parity(st & 0x6D);
            sr = (uint8_t)(((sr << 1) | in) & 0x3F);
        }
    }
    memset(out_bits, 0, ORA_ENCODED_BITS);
    for (size_t i = 0; i < ORA_ENCODED_BITS; ++i) { /* mod:142-153 */
        size_t p = (i % 32) * 67 + (i / 32);
        size_t corrected = (p / 8) * 8 + (7 - p % 8);
        out_bits[corrected] = enc[i];
    }
}

typedef struct { double ph1, ph2; int xor_T; int b_n; } mod_t;

static void modulate_bit(mod_t* m, int tx_bit, int16_t* out) { /* mod:228-284 */
    int d_val = (tx_bit == 0) ? 1 : -1;
    int d_val_xor;
    if (d_val == 1 && m->xor_T == 1) d_val_xor = 1;
    else if (d_val == 1 && m->xor_T == -1) d_val_xor = -1;
    else if (d_val == -1 && m->xor_T == 1) d_val_xor = -1;
    else if (d_val == -1 && m->xor_T == -1) d_val_xor = 1;
    else d_val_xor = 1;
    int d_pos = (d_val + 1) >> 1;
    int d_neg = (d_val - 1) >> 1;
    int d_pos_enc = d_pos;
    int d_neg_enc = (m->b_n == 0) ? d_neg : -d_neg;
    int d_s1, d_s2;
    if (d_pos_enc == 1 && m->xor_T == 1) d_s1 = 1;
    else if (d_pos_enc == 1 && m->xor_T == -1) d_s1 = -1;
    else d_s1 = 0;
    if (d_neg_enc == -1 && m->xor_T == 1) d_s2 = -1;
    else if (d_neg_enc == -1 && m->xor_T == -1) d_s2 = 1;
    else if (d_neg_enc == 1 && m->xor_T == 1) d_s2 = 1;
    else if (d_neg_enc == 1 && m->xor_T == -1) d_s2 = -1;
    else d_s2 = 0;
    double inc1 = TWO_PI * (-FREQ_DEV) / SAMPLE_RATE;
    double inc2 = TWO_PI * (+FREQ_DEV) / SAMPLE_RATE;
    for (size_t i = 0; i < ORA_SPS; ++i) {
        double s1 = sin(m->ph1), c1 = cos(m->ph1), s2 = sin(m->ph2), c2 = cos(m->ph2);
        double I = d_s1 * s1 + d_s2 * s2;
        double Q = d_s1 * c1 + d_s2 * c2;
        out[2 * i] = (int16_t)(16383.0 * I);
        out[2 * i + 1] = (int16_t)(16383.0 * Q);
        m->ph1 += inc1; m->ph2 += inc2;
        while (m->ph1 > PI) m->ph1 -= TWO_PI;
        while (m->ph1 < -PI) m->ph1 += TWO_PI;
        while (m->ph2 > PI) m->ph2 -= TWO_PI;
        while (m->ph2 < -PI) m->ph2 += TWO_PI;
    }
    m->xor_T = d_val_xor;
    m->b_n = 1 - m->b_n;
}

size_t ora_modulate_frames(const uint8_t* frames, size_t n_frames, int16_t* out) { /* mod:473-529 */
    mod_t m = {0.0, 0.0, 0, 1}; /* reset(): mod:221-226 */
    size_t o = 0;
    uint8_t bits[ORA_ENCODED_BITS];
    for (size_t f = 0; f < n_frames; ++f) {
        ora_encode_frame(frames + f * ORA_FRAME_BYTES, bits);
        for (int i = 23; i >= 0; --i) { modulate_bit(&m, (SYNC_WORD >> i) & 1, out + 2 * o); o += ORA_SPS; } /* mod:315-321 */
        for (size_t i = 0; i < ORA_ENCODED_BITS; ++i) { modulate_bit(&m, bits[i], out + 2 * o); o += ORA_SPS; }
    }
    memset(out + 2 * o, 0, sizeof(int16_t) * 2 * 100 * ORA_SPS); /* mod:527-529 */
    o += 100 * ORA_SPS;
    return o;
}

"""CPU checks of the counter-based dropout mask (vitb200/csrc/dropout.cuh is host-compilable): keep fraction, threshold
quantisation, decorrelation between sites / seeds, and known answers of the hash that pin the random stream the GPU tests replay."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <cstdio>
#include <cstdlib>
#include "dropout.cuh"
using namespace vb;
int main() {
    // known answers (murmur3 finaliser)
    printf("%u %u %u\n", fmix32(0u), fmix32(1u), fmix32(0xdeadbeefu));
    printf("%u %u %u\n", dropout_threshold(0.f), dropout_threshold(0.1f), dropout_threshold(0.5f));
    const float ps[3] = {0.1f, 0.25f, 0.5f};
    for (int t = 0; t < 3; ++t) {
        const uint32_t th = dropout_threshold(ps[t]);
        const uint32_t k1 = dropout_key(7u, 3u), k2 = dropout_key(7u, 4u), k3 = dropout_key(8u, 3u);
        long keep = 0, both12 = 0, both13 = 0, n = 1 << 22;
        for (uint32_t i = 0; i < (uint32_t)n; ++i) {
            const bool a = dropout_keep(k1, i, th), b = dropout_keep(k2, i, th), c = dropout_keep(k3, i, th);
            keep += a; both12 += a && b; both13 += a && c;
        }
        printf("%f %f %f\n", (double)keep / n, (double)both12 / n, (double)both13 / n);
    }
    return 0;
}
'''


def test_dropout_hash_statistics_and_known_answers(tmp_path):
    src = tmp_path / "d.cpp"
    src.write_text(SRC)
    exe = tmp_path / "d"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "vitb200", "csrc"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split("\n")
    assert out[0].split() == ["0", "1364076727", "233162409"]          # fmix32(0), fmix32(1), fmix32(0xdeadbeef)
    assert out[1].split() == ["0", "1677722", "8388608"]                # round(p * 2^24)
    for line, p in zip(out[2:5], (0.1, 0.25, 0.5)):
        keep, both_site, both_seed = map(float, line.split())
        q = 1.0 - p
        assert abs(keep - q) < 2e-3                                      # Bernoulli(1 - p) keep rate over 4 M elements
        assert abs(both_site - q * q) < 3e-3 and abs(both_seed - q * q) < 3e-3   # different site / seed: independent masks

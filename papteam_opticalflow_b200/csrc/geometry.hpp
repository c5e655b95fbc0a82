// Host-side pyramid geometry, evaluated in double exactly as the reference evaluates it, so that
// level sizes, Gaussian half-widths and resize ratios are bit-identical.
#pragma once
#include <cmath>
#include <vector>

namespace pf {

struct Level {
    int w, h;        // level size
    int src;         // level this one is built from (0 for the first n levels, i-n afterwards)
    int half;        // Gaussian half-width: (int)(sigma*3)
    double sigma;    // Gaussian sigma
    double rate;     // resize ratio applied to the blurred source
    std::vector<double> taps;  // normalised Gaussian taps, 2*half+1 entries
};

// S/GaussianPyramid.cpp:50-51: out-of-range ratios are silently replaced INSIDE the pyramid only.
inline double effective_ratio(double ratio) { return (ratio > 0.98 || ratio < 0.4) ? 0.75 : ratio; }

// S/GaussianPyramid.cpp:53: nLevels = log(minWidth/width)/log(ratio), truncated by the int store.
inline int levels_from_min_width(int width, double ratio, int min_width) {
    ratio = effective_ratio(ratio);
    return (int)(std::log((double)min_width / width) / std::log(ratio));
}

// S/GaussianPyramid.cpp:89-106 (sizes: S/Image.h:755-756; taps: S/Image.h:1210-1219).
inline std::vector<Level> level_geometry(int w0, int h0, double ratio, int nlevels) {
    ratio = effective_ratio(ratio);
    std::vector<Level> L((size_t)nlevels);
    if (nlevels <= 0) return L;
    double base_sigma = 1 / ratio - 1;
    int n = (int)(std::log(0.25) / std::log(ratio));
    double n_sigma = base_sigma * n;
    L[0] = Level{w0, h0, 0, 0, 0.0, 1.0, {1.0}};
    for (int i = 1; i < nlevels; i++) {
        Level& l = L[(size_t)i];
        if (i <= n) {
            l.src = 0;
            l.sigma = base_sigma * i;
            l.rate = std::pow(ratio, i);
        } else {
            l.src = i - n;
            l.sigma = n_sigma;
            l.rate = (double)std::pow(ratio, i) * w0 / L[(size_t)l.src].w;
        }
        l.half = (int)(l.sigma * 3);
        l.w = (int)((double)L[(size_t)l.src].w * l.rate);
        l.h = (int)((double)L[(size_t)l.src].h * l.rate);
        double two_s2 = l.sigma * l.sigma * 2, sum = 0;
        l.taps.resize((size_t)(2 * l.half + 1));
        for (int t = -l.half; t <= l.half; t++) {
            l.taps[(size_t)(t + l.half)] = std::exp(-(double)(t * t) / two_s2);
            sum += l.taps[(size_t)(t + l.half)];
        }
        for (auto& t : l.taps) t /= sum;
    }
    return L;
}

}  // namespace pf

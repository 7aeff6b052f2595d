// oracle/shim/ensmallen_bits/lbfgs/lbfgs.hpp — TEST INFRASTRUCTURE ONLY.
// Stand-in for ens::L_BFGS (ensmallen 2.x, un-vendored: vcpkg.json:8) so that core_private.cpp's
// call site (:264-294) compiles.  It is the same restatement of the published algorithm as
// oracle/rssync_oracle.cpp::lbfgs3 (defaults numBasis=10, armijoConstant=1e-4, wolfe=0.9,
// factr=1e-15, maxLineSearchTrials=50, minStep=1e-20, maxStep=1e20), written for a generic
// n-vector.  PARITY UNPINNED: the real ensmallen is not available to check against.
#pragma once
#include <armadillo>
#include <cmath>
#include <limits>
#include <vector>

namespace ens {

class L_BFGS {
   public:
    size_t& MaxIterations() { return maxIterations; }
    double& MinGradientNorm() { return minGradientNorm; }

    template <class FunctionType>
    double Optimize(FunctionType& function, arma::mat& iterate) {
        const size_t n = iterate.n_elem;
        const size_t numBasis = 10, maxTrials = 50;
        const double armijo = 1e-4, wolfe = 0.9, factr = 1e-15, minStep = 1e-20, maxStep = 1e20;
        std::vector<std::vector<double>> S(numBasis, std::vector<double>(n)), Y(numBasis, std::vector<double>(n));
        std::vector<double> g(n), oldx(n), oldg(n), dir(n), rho(numBasis), alpha(numBasis);
        arma::mat grad, trial(iterate.n_rows, iterate.n_cols);
        auto dotv = [n](const double* a, const double* b) {  // arma::dot, plain loop
            double s = 0.0;
            for (size_t i = 0; i < n; ++i) s += a[i] * b[i];
            return s;
        };
        double f = function.EvaluateWithGradient(iterate, grad);
        for (size_t c = 0; c < n; ++c) g[c] = grad[c];
        double* x = iterate.memptr();
        for (size_t it = 0; it != maxIterations; ++it) {
            const double prevf = f;
            if (it > 0 && std::sqrt(dotv(g.data(), g.data())) < minGradientNorm) break;
            if (std::isnan(f)) break;
            double scaling;
            if (it > 0) {
                const size_t pp = (it - 1) % numBasis;
                const double yy = dotv(Y[pp].data(), Y[pp].data());
                const double denom = (yy >= 1e-10) ? yy : 1.0;
                scaling = dotv(S[pp].data(), Y[pp].data()) / denom;
            } else {
                const double gn = std::sqrt(dotv(g.data(), g.data()));
                scaling = (gn >= 1e-5) ? 1.0 / gn : 1.0;
            }
            if (scaling == 0.0 || !std::isfinite(scaling)) break;
            dir = g;
            const size_t limit = (numBasis > it) ? 0 : (it - numBasis);
            for (size_t i = it; i != limit; --i) {
                const size_t tp = (i + (numBasis - 1)) % numBasis;
                const double ys = dotv(Y[tp].data(), S[tp].data());
                rho[it - i] = (ys != 0) ? (1.0 / ys) : 1.0;
                alpha[it - i] = rho[it - i] * dotv(S[tp].data(), dir.data());
                for (size_t c = 0; c < n; ++c) dir[c] -= alpha[it - i] * Y[tp][c];
            }
            for (size_t c = 0; c < n; ++c) dir[c] *= scaling;
            for (size_t i = limit; i < it; ++i) {
                const size_t tp = i % numBasis;
                const double beta = rho[it - i - 1] * dotv(Y[tp].data(), dir.data());
                const double coef = alpha[it - i - 1] - beta;
                for (size_t c = 0; c < n; ++c) dir[c] += coef * S[tp][c];
            }
            for (size_t c = 0; c < n; ++c) dir[c] = -dir[c];
            for (size_t c = 0; c < n; ++c) { oldx[c] = x[c]; oldg[c] = g[c]; }
            double step = 1.0, bestStep = 1.0, bestObj = std::numeric_limits<double>::max();
            const double init_dg = dotv(g.data(), dir.data());
            if (init_dg > 0.0) break;
            const double f0 = f;
            const double lin = armijo * init_dg;
            size_t trials = 0;
            for (;;) {
                for (size_t c = 0; c < n; ++c) trial[c] = x[c] + step * dir[c];
                f = function.EvaluateWithGradient(trial, grad);
                for (size_t c = 0; c < n; ++c) g[c] = grad[c];
                if (f < bestObj) { bestStep = step; bestObj = f; }
                trials++;
                double width;
                if (f > f0 + step * lin) {
                    width = 0.5;
                } else {
                    const double dg = dotv(g.data(), dir.data());
                    if (dg < wolfe * init_dg) width = 2.1;
                    else if (dg > -wolfe * init_dg) width = 0.5;
                    else break;
                }
                if (step < minStep || step > maxStep || trials >= maxTrials) break;
                step *= width;
            }
            for (size_t c = 0; c < n; ++c) x[c] += bestStep * dir[c];
            if (bestStep == 0.0) break;
            const double denom = std::max(std::max(std::fabs(prevf), std::fabs(f)), 1.0);
            if ((prevf - f) / denom <= factr) break;
            const size_t op = it % numBasis;
            for (size_t c = 0; c < n; ++c) { S[op][c] = x[c] - oldx[c]; Y[op][c] = g[c] - oldg[c]; }
        }
        return f;
    }

   private:
    size_t maxIterations = 10000;
    double minGradientNorm = 1e-6;
};

}  // namespace ens

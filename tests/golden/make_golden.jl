# Reference-side golden generator: runs the REAL GPCC.jl (HITS-AIN/GPCC.jl v0.1.35) on the inputs of the committed
# fixtures and writes the reference's outputs under the same keys, so that the oracle and the CUDA library can be pinned
# against the reference itself.  The reference ships no tests or vectors (test/runtests.jl:4-6) and Julia is not
# available in the build image, so this script could not be executed there; a maintainer runs it once:
#
#   julia -e 'using Pkg; Pkg.Registry.add(RegistrySpec(url="https://github.com/HITS-AIN/AINJuliaRegistry")); Pkg.add(["GPCC","MiscUtil","NPZ","Distributions"])'
#   julia tests/golden/make_golden.jl            # from the repository root; writes tests/golden/julia_*.npz
#   python -m pytest tests/test_julia_golden.py  # compares oracle (CPU) and, with -m gpu, the CUDA library with them
#
# What is pinned, with the reference code that produces it:
#   julia_miscutil.npz    MiscUtil.makepositive / invmakepositive / transformbetween / invtransformbetween on probe points
#                         (the un-vendored transforms behind unpack, src/gpccfixdelay_marginaliseb.jl:112-126, :195-196)
#   julia_loglik_*.npz    for every loglik_*.npz fixture: K = GPCC.delayedCovariance(...) (src/delayedCovariance.jl:1-38) and
#                         logpdf(MvNormal(bbar, K + Sobs + B), Y) exactly as the objective builds it (:85-98, :133-141)
#   julia_fit_cfg1.npz    gpcc(...) on the cfg1 data (:46-53): loglikel, alpha, rho, postb, pred (three methods), together
#                         with the start points the reference drew from MersenneTwister(seed) (:62, :166, :188, :195-196),
#                         re-drawn here in the same order so that the library can be started from identical theta0
#   julia_grid_cfg2.npz   the README's 1-D grid (README.md:170-178) + getprobabilities with and without the delay prior
using GPCC, MiscUtil, NPZ, Distributions, LinearAlgebra, Random, Printf, Statistics

const HERE = @__DIR__
const KERNELS = Dict("OU" => GPCC.OU, "rbf" => GPCC.rbf, "matern32" => GPCC.matern32, "matern52" => GPCC.matern52)

splitbands(v, n) = [v[(sum(n[1:l-1])+1):sum(n[1:l])] for l in 1:length(n)]

# ---- MiscUtil probes ---------------------------------------------------------------------------------------------
let x = collect(range(-12.0, 12.0, length = 49)), lo = 0.1, hi = 300.0
    pos = MiscUtil.makepositive.(x)
    btw = [MiscUtil.transformbetween(xi, lo, hi) for xi in x]
    npzwrite(joinpath(HERE, "julia_miscutil.npz"), Dict(
        "x" => x, "makepositive" => pos, "invmakepositive" => MiscUtil.invmakepositive.(pos),
        "lo" => lo, "hi" => hi, "transformbetween" => btw,
        "invtransformbetween" => [MiscUtil.invtransformbetween(b, lo, hi) for b in btw]))
end

# ---- fixed hyper-parameter log-likelihood, the objective of :133-141 restated around the reference's own functions ----
function reference_loglik(tarray, yarray, stdarray, kernel, delays, alpha, rho)
    Y    = reduce(vcat, yarray)                                   # :85
    Q    = GPCC.Qmatrix(length.(tarray))                          # :87   (name per src/util.jl:56-70; adapt if it differs)
    Sobs = Diagonal(reduce(vcat, stdarray) .^ 2)                  # :89
    mub  = map(mean, yarray)                                      # :92
    Sigb = 100.0 * diagm(map(var, yarray))                        # :94
    B    = Q * Sigb * Q'                                          # :96
    bbar = Q * mub                                                # :98
    K    = GPCC.delayedCovariance(kernel, alpha, delays, rho, tarray) + Sobs + B      # :135
    K    = Matrix(Symmetric(K))                                   # makematrixsymmetric! (:137)
    return logpdf(MvNormal(bbar, K), Y), mub, diag(Sigb)           # :139
end

for f in filter(x -> startswith(x, "loglik_") && endswith(x, ".npz") && x != "loglik_large.npz", readdir(HERE))
    g = npzread(joinpath(HERE, f))
    n = Int.(g["n"])
    t, y, s = splitbands(g["t"], n), splitbands(g["y"], n), splitbands(g["s"], n)
    kern = KERNELS[String(g["kernel"])]                           # stored as a 0-d string array; adapt the decoding if NPZ differs
    M = size(g["delays"], 1)
    ll = zeros(M)
    mub, Sigb = zeros(length(n)), zeros(length(n))
    for m in 1:M
        ll[m], mub, Sigb = reference_loglik(t, y, s, kern, g["delays"][m, :], g["alpha"][m, :], g["rho"][m])
    end
    npzwrite(joinpath(HERE, "julia_" * f), Dict("loglik" => ll, "mub" => mub, "Sigmab" => Sigb))
    @printf("%s: %d log-likelihoods\n", f, M)
end

# ---- start points as the reference draws them (:62, :166, :188, :195-196), same RNG, same order ---------------------
function reference_theta0(yarray; seed = 1, numberofrestarts = 1, initialrandom = 5, rhomin = 0.1, rhomax = 20.0)
    rg = MersenneTwister(seed)                                    # :62
    L = length(yarray)
    rho0 = numberofrestarts <= 2 ? rand(rg, Uniform(rhomin + 1e-3, rhomax - 1e-3), numberofrestarts) :    # :166
                                   collect(MiscUtil.logrange(rhomin + 1e-3, rhomax - 1e-3, numberofrestarts))  # :172
    th = zeros(numberofrestarts, initialrandom, L + 1)
    for i in 1:numberofrestarts, j in 1:initialrandom             # getsolution(i) draws `initialrandom` solutions (:207)
        a = map(var, yarray) .* (rand(rg, L) * (1.2 - 0.8) .+ 0.8)        # :188
        th[i, j, :] = [MiscUtil.invmakepositive.(a); MiscUtil.invtransformbetween(rho0[i], rhomin, rhomax)]   # :195-196
    end
    return th, rho0
end

# ---- cfg1: single fit on the data of fit_cfg1_cfg2.npz ---------------------------------------------------------------
let g = npzread(joinpath(HERE, "fit_cfg1_cfg2.npz"))
    n = Int.(g["n"])
    t, y, s = splitbands(g["t"], n), splitbands(g["y"], n), splitbands(g["s"], n)
    delays = g["truedelays"]
    th, rho0 = reference_theta0(y; seed = 1, initialrandom = 5, rhomin = 0.1, rhomax = 300.0)
    loglikel, pred, (alpha, postb, rho) = gpcc(t, y, s; kernel = GPCC.matern32, delays = delays, iterations = 1000, rhomax = 300)
    ttest = g["ttest"]
    mu, sd = pred(ttest)                                          # :293-307
    mufull, Sfull = pred([ttest[1:7], ttest[6:9]])                # :259-289 (the ragged test sets of the Python fixture)
    tl = pred([[9.0, 10.0, 11.0], [9.0, 10.0, 11.0]], [[6.34, 5.49, 5.38], [13.08, 12.37, 15.69]], [[0.34, 0.42, 0.2], [0.87, 0.8, 0.66]])   # :311-343
    npzwrite(joinpath(HERE, "julia_fit_cfg1.npz"), Dict(
        "theta0" => th[1, :, :], "rho0" => rho0, "loglikel" => loglikel, "alpha" => alpha, "rho" => rho,
        "postb_mu" => mean(postb), "postb_Sigma" => Matrix(cov(postb)),
        "pred_mu" => reduce(hcat, mu)', "pred_sd" => reduce(hcat, sd)', "predfull_mu" => mufull, "predfull_Sigma" => Matrix(Sfull),
        "test_loglik" => tl))
    @printf("cfg1: loglikel %.10f\n", loglikel)
end

# ---- cfg2: the README's 1-D grid and getprobabilities ----------------------------------------------------------------
let g = npzread(joinpath(HERE, "fit_cfg1_cfg2.npz"))
    n = Int.(g["n"])
    t, y, s = splitbands(g["t"], n), splitbands(g["y"], n), splitbands(g["s"], n)
    cands = g["cands"]
    ll = map(d -> gpcc(t, y, s; kernel = GPCC.matern32, delays = [0.0; d], iterations = 1000, rhomax = 300)[1], cands)   # README.md:172-174
    prior = uniformpriordelay(; L = 1e44, z = 0.0)                # src/uniformpriordelay.jl:10-16
    npzwrite(joinpath(HERE, "julia_grid_cfg2.npz"), Dict(
        "cands" => cands, "ll_grid" => ll, "post_flat" => getprobabilities(ll),
        "post_prior" => getprobabilities(ll, logpdf.(prior, cands)), "prior_upper" => maximum(prior)))
    @printf("cfg2: grid mode at %.1f\n", cands[argmax(ll)])
end

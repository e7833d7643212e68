# export_golden.jl -- dump golden tables and values from a REAL ACE.jl (v0.12.46) installation.
#
# Run on any machine that has Julia + ACE.jl (the image this repository is built in has neither):
#
#     julia --project=<env with ACE, JSON, StaticArrays> tests/golden/export_golden.jl [outdir]
#
# It writes one JSON file per configuration into `outdir` (default: tests/golden/julia/).  Commit those files:
# tests/test_julia_golden.py picks up every tests/golden/julia/*.json and checks the CPU oracle (and, on a GPU
# box, the CUDA path through the C ABI) against the values ACE.jl itself produced -- from that moment the
# files supersede the oracle as the parity reference (SURVEY.md, Appendix C.7).
#
# Schema "aceb200-golden-v1" (every floating-point number is printed with Julia's shortest round-trip
# representation; complex values are interleaved (re, im); arrays of StaticArrays are flattened column-major,
# i.e. exactly in Julia's memory order, which is the order the C ABI uses):
#
#   rn        pl, tl, pr, tr, A, B, C            OrthPolyBasis fields            (src/polynomials/orthpolys.jl:82-92)
#             trans_exstr                          Lambda.exstr                    (src/transforms/lambdas.jl:9-12)
#   maxL                                           SHBasis degree                  (src/b1pcomponents/Ylm.jl:18-26)
#   comp_kinds, categories                         Product1pBasis.bases in order   (src/product_1pbasis.jl:5-8)
#   indices   [nA][NB]                             Product1pBasis.indices, 1-based
#   spec1p    [nA] dictionaries                    get_spec(basis1p)
#   orders, iAA2iA [nAA][maxord]                   PIBasisSpec                     (src/pibasis.jl:10-13)
#   pireal, symreal, property                      pibasis.real, basis.real, typeof(phi)
#   A2B       m, n, I, J (1-based), V [nnz][ncomp] complex   findnz(A2Bmap)        (src/symmbasis.jl:33-38)
#   c         [nB][nprop]                          LinearACEModel.c
#   ctilde    [nAA][nprop][ncomp] complex          ProductEvaluator.coeffs         (src/evaluator.jl:11)
#   envs[k]   R [J][3], species [J] (1-based category index) or null,
#             A [nA] complex, AA [nAA] complex, B [nB][ncomp] complex,
#             dA [J][nA][3] complex, dAA [J][nAA][3] complex, dB [J][nB][3][ncomp] complex,
#             E [nprop][ncomp] complex, G [J][nprop][3][ncomp] complex
#
# Values that are real in Julia are written with a zero imaginary part so that one reader handles all cases.

using ACE, JSON, StaticArrays, Random, LinearAlgebra, SparseArrays
using ACE: evaluate, evaluate_d, evaluate_ed, grad_config, get_spec, PositionState, ACEConfig, State

outdir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "julia")
mkpath(outdir)

# ---- flatten anything on the path to interleaved (re, im) Float64, column-major ------------------------
cflat(x::Number) = Float64[real(x), imag(x)]
cflat(x::ACE.AbstractProperty) = cflat(x.val)
cflat(x::ACE.DState) = cflat(x.rr)
cflat(x::AbstractArray) = isempty(x) ? Float64[] : reduce(vcat, [cflat(v) for v in vec(collect(x))])

propname(φ) = string(nameof(typeof(φ)))

function component_kind(b)
   if b isa ACE.Categorical1pBasis
      return "Cat"
   end
   syms = ACE.symbols(b)
   return length(syms) == 1 ? "Rn" : "Ylm"
end

function rn_tables(Rn)
   # Rn.basis = chain(norm, trans, OrthPolyBasis)  (src/b1pcomponents/Rn.jl:21)
   trans = Rn.basis.F[2]
   J = Rn.basis.F[3]
   return Dict("pl" => J.pl, "tl" => J.tl, "pr" => J.pr, "tr" => J.tr,
               "A" => J.A, "B" => J.B, "C" => J.C, "trans_exstr" => trans.exstr)
end

function dump_config(name, φ, B1p, Bsel; nprop = 1, J = 10, nenvs = 3, seed = 20240, categories = nothing)
   Random.seed!(seed)
   basis = ACE.SymmetricBasis(φ, B1p, ACE.O3(), Bsel)      # src/symmbasis.jl:74-82
   b1p = basis.pibasis.basis1p
   nB = length(basis)
   W = rand(nprop, nB) .- 0.5
   c = nprop == 1 ? W[1, :] : [SVector{nprop}(W[:, i]) for i = 1:nB]
   model = ACE.LinearACEModel(basis, c, evaluator = :standard)

   kinds = [component_kind(b) for b in b1p.bases]
   iRn = findfirst(==("Rn"), kinds)
   iY = findfirst(==("Ylm"), kinds)
   I, Jc, V = findnz(basis.A2Bmap)
   D = Dict{String, Any}(
      "schema" => "aceb200-golden-v1",
      "generator" => "ACE.jl " * string(pkgversion(ACE)),
      "config" => name,
      "rn" => rn_tables(b1p.bases[iRn]),
      "maxL" => b1p.bases[iY].basis.alp.L,                    # SHBasis.alp::ALPolynomials (sphericalharmonics.jl:125, 285)
      "comp_kinds" => kinds,
      "categories" => categories === nothing ? nothing : string.(categories),
      "indices" => [collect(t) for t in b1p.indices],
      "spec1p" => [Dict(string(k) => (v isa Symbol ? string(v) : v) for (k, v) in pairs(b)) for b in get_spec(b1p)],
      "orders" => basis.pibasis.spec.orders,
      "iAA2iA" => [basis.pibasis.spec.iAA2iA[i, :] for i = 1:length(basis.pibasis)],
      "pireal" => basis.pibasis.real === Base.real,
      "symreal" => basis.real === Base.real,
      "property" => propname(φ),
      "A2B" => Dict("m" => size(basis.A2Bmap, 1), "n" => size(basis.A2Bmap, 2), "I" => I, "J" => Jc,
                    "V" => [cflat(v) for v in V]),
      "nprop" => nprop,
      "c" => [W[:, i] for i = 1:nB],
      "ctilde" => [cflat(v) for v in model.evaluator.coeffs],
   )
   envs = Any[]
   Rn = b1p.bases[iRn]
   for k = 1:nenvs
      Rs = [ACE.rand_radial(Rn) * ACE.rand_sphere() for _ = 1:J]      # src/utils/random.jl:22-25
      if categories === nothing
         Xs = [PositionState(r) for r in Rs]
         species = nothing
      else
         sp = rand(1:length(categories), J)
         Xs = [State(rr = Rs[j], mu = categories[sp[j]]) for j = 1:J]
         species = sp
      end
      cfg = ACEConfig(Xs)
      A, dA = evaluate_ed(b1p, cfg)
      AA, dAA = evaluate_ed(basis.pibasis, cfg)
      B, dB = evaluate_ed(basis, cfg)
      E = evaluate(model, cfg)
      G = grad_config(model, cfg)
      # Matrix{DState}(nbasis x J) is column-major: memory order [neighbour][basis index] -- what the C ABI uses
      push!(envs, Dict(
         "R" => [collect(r) for r in Rs], "species" => species,
         "A" => cflat(A), "AA" => cflat(AA), "B" => cflat(B),
         "dA" => cflat(dA), "dAA" => cflat(dAA), "dB" => cflat(dB),
         "E" => cflat(E), "G" => cflat(G)))
   end
   D["envs"] = envs
   open(joinpath(outdir, name * ".json"), "w") do io
      JSON.print(io, D)
   end
   @info "wrote $(name): nA = $(length(b1p)), nAA = $(length(basis.pibasis)), nB = $(nB)"
end

sparse_sel(ord, deg; wL = 1.5) = ACE.SparseBasis(; maxorder = ord, p = 1, default_maxdeg = deg,
                                                 weight = Dict(:n => 1.0, :l => wL))
rnylm(deg, Bsel; wL = 1.5) = ACE.Utils.RnYlm_1pbasis(maxdeg = deg, maxL = ceil(Int, deg / wL), Bsel = Bsel)

# the reference's unit-test basis (test/test_symmbasis.jl, test/test_linearmodel.jl)
let Bsel = ACE.SimpleSparseBasis(3, 6)
   dump_config("inv_simple_3_6", ACE.Invariant(), ACE.Utils.RnYlm_1pbasis(maxdeg = 6), Bsel)
end
# BASELINE config 1 (benchmark/bm_basis.jl:58-62 with maxorder = ord) and config 2 (benchmark/bm_linear.jl:151-155)
let Bsel = sparse_sel(3, 10); dump_config("config1_inv_sparse_3_10", ACE.Invariant(), rnylm(10, Bsel), Bsel; J = 30) end
let Bsel = sparse_sel(3, 12); dump_config("config2_inv_sparse_3_12", ACE.Invariant(), rnylm(12, Bsel), Bsel; J = 40) end
# BASELINE config 3 (profile/profile_linearmodel.jl:13-25): tables only matter here (round-off-dependent cleaning)
let Bsel = sparse_sel(4, 14); dump_config("config3_inv_sparse_4_14", ACE.Invariant(), rnylm(14, Bsel), Bsel; J = 60, nenvs = 1) end
# BASELINE config 4: equivariant bases
let Bsel = sparse_sel(3, 10); dump_config("config4_euclvec_3_10", ACE.EuclideanVector(Float64), rnylm(10, Bsel), Bsel; J = 30, nenvs = 1) end
let Bsel = sparse_sel(3, 10); dump_config("config4_euclmat_3_10", ACE.EuclideanMatrix(Float64), rnylm(10, Bsel), Bsel; J = 30, nenvs = 1) end
# BASELINE config 5's family at test size (test/test_discrete.jl:67-87): 4 species, multi-property
let Bsel = sparse_sel(3, 5)
   cats = [:a, :b, :c, :d]
   B1p = ACE.Categorical1pBasis(cats; varsym = :mu, idxsym = :q) * ACE.Utils.RnYlm_1pbasis(maxdeg = 5, maxL = 4)
   dump_config("config5_species_3_5", ACE.Invariant(), B1p, Bsel; nprop = 4, J = 12, categories = cats)
end

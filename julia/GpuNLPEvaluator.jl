# GpuNLPEvaluator.jl -- the binding a Powersense.jl maintainer adds to use libpowersense_b200.so.
# UNEXECUTED in the authoring container (no Julia there); written against include/powersense_b200.h.
#
# It subtypes MOI.AbstractNLPEvaluator and implements exactly the method set of the reference's own
# EmptyNLPEvaluator (src/opt/MOI_wrapper.jl:56-77), forwarding each call to the C ABI with `ccall`.
module PowersenseB200

import MathOptInterface
const MOI = MathOptInterface

const LIB = get(ENV, "POWERSENSE_B200_LIB", "libpowersense_b200.so")

# struct ps_network (include/powersense_b200.h); Julia vectors are passed by pointer and copied at ps_create
struct PsNetwork
    nbus::Int64; nbr::Int64; ngen::Int64; nbss::Int64; n_tgh::Int64
    br_f::Ptr{Int64}; br_t::Ptr{Int64}
    Gf::Ptr{Float64}; Bf::Ptr{Float64}; Gt::Ptr{Float64}; Bt::Ptr{Float64}
    gf::Ptr{Float64}; bf::Ptr{Float64}; gt::Ptr{Float64}; bt::Ptr{Float64}
    Imax::Ptr{Float64}
    Pd::Ptr{Float64}; Qd::Ptr{Float64}; Gl::Ptr{Float64}; Bl::Ptr{Float64}
    gen_ptr::Ptr{Int64}; gen_idx::Ptr{Int64}; sij_ptr::Ptr{Int64}; sij_idx::Ptr{Int64}
    sji_ptr::Ptr{Int64}; sji_idx::Ptr{Int64}; sh_ptr::Ptr{Int64}; sh_idx::Ptr{Int64}
    c0::Ptr{Float64}; c1::Ptr{Float64}; c2::Ptr{Float64}
    cost_order::Int32
end

check(h, rc) = rc == 0 || error("powersense_b200: " * unsafe_string(ccall((:ps_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))

# formulation codes follow src/opf/type.jl:7-15
const FORMULATION_CODE = Dict(:PBRAPVmodel => 0, :PBRARVmodel => 1, :PBRAWVmodel => 2, :CBRARVmodel => 3, :CBRAWVmodel => 4,
                              :PNPAPVmodel => 5, :PNRAPVmodel => 6, :PNRARVmodel => 7, :PNRAWVmodel => 8)

mutable struct GpuNLPEvaluator <: MOI.AbstractNLPEvaluator
    handle::Ptr{Cvoid}
    n_var::Int; m_nl::Int; nnz_jac::Int; nnz_hess::Int
    enable_hessian::Bool
end

"CSR lists from data.node[string(i)] = Dict(\"gen\"=>..., \"Sij\"=>..., \"Sji\"=>..., \"shunt\"=>...) (PowersenseData.jl:212)"
function csr(data, key)
    ptr = Int64[0]; idx = Int64[]
    for i in 1:data.nbus
        append!(idx, Int64.(data.node[string(i)][key]))
        push!(ptr, length(idx))
    end
    return ptr, idx
end

function GpuNLPEvaluator(data, formulation::Symbol; FACTS_bShunt = true, obj_type = :linear, device = 0)
    br_f = Int64[b[1] for b in data.br]; br_t = Int64[b[2] for b in data.br]
    gp, gi = csr(data, "gen"); ijp, iji = csr(data, "Sij"); jip, jii = csr(data, "Sji"); shp, shi = csr(data, "shunt")
    flags = (FACTS_bShunt ? 1 : 0) | (obj_type == :linear ? 2 : 0)
    src = data.source_type == "matpower" ? 0 : 1
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve br_f br_t gp gi ijp iji jip jii shp shi data begin
        net = PsNetwork(data.nbus, data.nbr, data.ngen, data.nbss, sum(data.lps),
                        pointer(br_f), pointer(br_t),
                        pointer(data.Gf), pointer(data.Bf), pointer(data.Gt), pointer(data.Bt),
                        pointer(data.gf), pointer(data.bf), pointer(data.gt), pointer(data.bt), pointer(data.Imax),
                        pointer(data.Pd), pointer(data.Qd), pointer(data.Gl), pointer(data.Bl),
                        pointer(gp), pointer(gi), pointer(ijp), pointer(iji), pointer(jip), pointer(jii), pointer(shp), pointer(shi),
                        pointer(data.c0), pointer(data.c1), pointer(data.c2), Int32(data.cost_order))
        rc = ccall((:ps_create, LIB), Cint, (Ref{PsNetwork}, Cint, Cint, Cint, Cint, Ref{Ptr{Cvoid}}),
                   net, FORMULATION_CODE[formulation], src, flags, device, out)
        rc == 0 || error("powersense_b200: " * unsafe_string(ccall((:ps_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    end
    n = Ref{Int64}(); m = Ref{Int64}(); nj = Ref{Int64}(); nh = Ref{Int64}()
    check(out[], ccall((:ps_dims, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Int64}), out[], n, m, nj, nh))
    ev = GpuNLPEvaluator(out[], n[], m[], nj[], nh[], false)
    finalizer(e -> ccall((:ps_destroy, LIB), Cvoid, (Ptr{Cvoid},), e.handle), ev)
    return ev
end

# ---- the MOI method set (src/opt/MOI_wrapper.jl:56-77) --------------------------------------------------------
MOI.features_available(::GpuNLPEvaluator) = [:Grad, :Jac, :Hess]

function MOI.initialize(d::GpuNLPEvaluator, features::Vector{Symbol})
    for f in features
        f in (:Grad, :Jac, :Hess) || error("Unsupported feature $f")
    end
    d.enable_hessian = :Hess in features
    return
end

MOI.eval_objective(::GpuNLPEvaluator, x) = error("the NLP block has no objective (has_objective = false)")
MOI.eval_objective_gradient(::GpuNLPEvaluator, g, x) = error("the NLP block has no objective (has_objective = false)")

# g / J / H may be contiguous SubArrays into the solver's arrays (MOI_wrapper.jl:967, 1025, 1059): pass their pointers
function MOI.eval_constraint(d::GpuNLPEvaluator, g, x)
    GC.@preserve g x check(d.handle, ccall((:ps_eval_constraint, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                                            d.handle, pointer(x), pointer(g)))
    return
end

function MOI.jacobian_structure(d::GpuNLPEvaluator)
    r = Vector{Int64}(undef, d.nnz_jac); c = Vector{Int64}(undef, d.nnz_jac)
    check(d.handle, ccall((:ps_jacobian_structure, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), d.handle, r, c))
    return Tuple{Int64,Int64}[(r[k], c[k]) for k in 1:d.nnz_jac]
end

function MOI.eval_constraint_jacobian(d::GpuNLPEvaluator, J, x)
    GC.@preserve J x check(d.handle, ccall((:ps_eval_jacobian, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                                            d.handle, pointer(x), pointer(J)))
    return
end

function MOI.hessian_lagrangian_structure(d::GpuNLPEvaluator)
    d.enable_hessian || error("Hessian requested but not initialized")
    r = Vector{Int64}(undef, d.nnz_hess); c = Vector{Int64}(undef, d.nnz_hess)
    check(d.handle, ccall((:ps_hessian_structure, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), d.handle, r, c))
    return Tuple{Int64,Int64}[(r[k], c[k]) for k in 1:d.nnz_hess]
end

function MOI.eval_hessian_lagrangian(d::GpuNLPEvaluator, H, x, σ, μ)
    d.enable_hessian || error("Hessian requested but not initialized")
    GC.@preserve H x μ check(d.handle, ccall((:ps_eval_hessian, LIB), Cint,
                                              (Ptr{Cvoid}, Ptr{Float64}, Cdouble, Ptr{Float64}, Ptr{Float64}),
                                              d.handle, pointer(x), σ, pointer(μ), pointer(H)))
    return
end

"""
    adopt_hessian_order!(d, jump_evaluator)

Make the library report its Hessian structure and values in exactly the order of a JuMP evaluator of the same model
(`JuMP.NLPEvaluator(model)` of the reference's own `create_opf_model!`, initialised with `:Hess`): JuMP's intra-block
off-diagonal order depends on its colouring / `Set` iteration order and cannot be derived outside Julia, so the order is
IMPORTED (ps_match_hessian_structure: matched per NL-row block as sets of variable pairs, either triangle).  After this
call `MOI.hessian_lagrangian_structure(d)` equals JuMP's pair for pair up to the triangle convention, and
`eval_hessian_lagrangian` fills `H` in that order.  The kernels are untouched (a gather applies the permutation).
"""
function adopt_hessian_order!(d::GpuNLPEvaluator, jump_evaluator)
    hs = MOI.hessian_lagrangian_structure(jump_evaluator)
    length(hs) == d.nnz_hess || error("Hessian structures differ in length: $(length(hs)) vs $(d.nnz_hess)")
    r = Int64[t[1] for t in hs]; c = Int64[t[2] for t in hs]
    check(d.handle, ccall((:ps_match_hessian_structure, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), d.handle, r, c))
    return d
end

"""
    pin!(arrays...) / unpin!(arrays...)

Page-lock vectors the solver owns and reuses at every iteration (its constraint, Jacobian-value and Hessian-value arrays):
the callbacks then copy into them at the link's full rate instead of through the driver's staging of pageable memory
(ps_host_register; on case13659pegase `g + J + H` are 12.6 MB per call: 0.93 -> 0.49 ms).  Optional; `unpin!` before the
arrays are freed or resized.
"""
function pin!(arrays::Vector{Float64}...)
    for a in arrays
        rc = ccall((:ps_host_register, LIB), Cint, (Ptr{Cvoid}, Int64), pointer(a), sizeof(a))
        rc == 0 || error("powersense_b200: " * unsafe_string(ccall((:ps_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    end
end
unpin!(arrays::Vector{Float64}...) = foreach(a -> ccall((:ps_host_unregister, LIB), Cint, (Ptr{Cvoid},), pointer(a)), arrays)

"NLPBlockData for MOI.set(backend, MOI.NLPBlock(), ...): bounds (0,0) for == rows, (-Inf,0) for <= rows"
function nlp_block(d::GpuNLPEvaluator)
    lo = Vector{Float64}(undef, d.m_nl); hi = Vector{Float64}(undef, d.m_nl)
    check(d.handle, ccall((:ps_constraint_bounds, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), d.handle, lo, hi))
    return MOI.NLPBlockData([MOI.NLPBoundsPair(lo[i], hi[i]) for i in 1:d.m_nl], d, false)
end

end # module

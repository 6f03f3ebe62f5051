# make_golden.jl -- dump the REFERENCE evaluator's outputs so that parity with libpowersense_b200 can be pinned.
#
# UNEXECUTED in the authoring container (no Julia there).  Run it where the reference's own environment resolves
# (Powersense.jl v0.0.6 with its Manifest: JuMP 0.21.10, MathOptInterface 0.9.22, PowerModels 0.18.3):
#
#     python tools/compare_julia_golden.py export  julia_golden/            # writes the evaluation points (x, mu) as raw Float64
#     julia --project=/path/to/Powersense.jl julia/make_golden.jl /path/to/Powersense.jl julia_golden/
#     python tools/compare_julia_golden.py compare julia_golden/            # report + Hessian-order permutations
#
# For every case the reference ships (examples/opf/14bus, examples/opf/case300.m, examples/opt/acopf/case3.m) and every AC
# formulation (src/opf/type.jl:7-15) it builds the model with the UNMODIFIED create_opf_model! (src/opf/formulations.jl),
# wraps it in JuMP's own NLPEvaluator -- the object Powersense.Optimizer and Ipopt call at src/opt/MOI_wrapper.jl:813, 855,
# 969, 1026, 1061 -- and writes, per <case>_<formulation>:
#     .g.bin .J.bin .H.bin .lo.bin .hi.bin   Float64      eval_constraint / eval_constraint_jacobian / eval_hessian_lagrangian, bounds
#     .jrow.bin .jcol.bin .hrow.bin .hcol.bin Int64       jacobian_structure / hessian_lagrangian_structure (1-based tuples)
#     .meta.txt                                           sizes and package versions
# and, per <case>, the PowersenseData tables create_PowersenseData produced (<case>_data.<field>.bin), so that the
# comparison can feed the library the reference's OWN tables (evaluator parity independent of reader parity) and check
# the repo's readers against them as a separate line.
# Raw little-endian arrays: no package beyond the reference's own Manifest is needed.
import Pkg
import JuMP
import MathOptInterface
const MOI = MathOptInterface

length(ARGS) >= 2 || error("usage: julia make_golden.jl <Powersense.jl checkout> <directory with the .x.bin / .mu.bin points>")
const REF = ARGS[1]
const DIR = ARGS[2]

using Powersense

const CASES = [
    ("case14", [joinpath(REF, "examples/opf/14bus/case.raw"), joinpath(REF, "examples/opf/14bus/case.rop")]),
    ("case300", joinpath(REF, "examples/opf/case300.m")),
    ("case3", joinpath(REF, "examples/opt/acopf/case3.m")),
]
# the nine singletons of src/opf/type.jl:7-15, in that order (= the PS_* formulation codes)
const FORMS = [
    ("PBRAPVmodel", Powersense.PBRAPVmodel), ("PBRARVmodel", Powersense.PBRARVmodel), ("PBRAWVmodel", Powersense.PBRAWVmodel),
    ("CBRARVmodel", Powersense.CBRARVmodel), ("CBRAWVmodel", Powersense.CBRAWVmodel), ("PNPAPVmodel", Powersense.PNPAPVmodel),
    ("PNRAPVmodel", Powersense.PNRAPVmodel), ("PNRARVmodel", Powersense.PNRARVmodel), ("PNRAWVmodel", Powersense.PNRAWVmodel),
]

readf64(path) = collect(reinterpret(Float64, read(path)))
function writebin(path, v::Vector{T}) where {T}
    open(path, "w") do io
        write(io, v)
    end
end

function pkgversion_of(name)
    for (_, info) in Pkg.dependencies()
        info.name == name && return string(info.version)
    end
    return "?"
end

"CSR lists of data.node[string(i)][key] (PowersenseData.jl:212): 0-based offsets, 1-based element ids"
function csr(data, key)
    ptr = Int64[0]; idx = Int64[]
    for i in 1:data.nbus
        append!(idx, Int64.(data.node[string(i)][key]))
        push!(ptr, length(idx))
    end
    return ptr, idx
end

function dump_data(cname, data)
    stem = joinpath(DIR, "$(cname)_data")
    for f in (:Vmax, :Vmin, :Pmax, :Pmin, :Qmax, :Qmin, :bmax, :bmin, :Imax, :Gf, :Bf, :Gt, :Bt, :gf, :bf, :gt, :bt,
              :Pd, :Qd, :Gl, :Bl, :c0, :c1, :c2)
        writebin(stem * ".$(f).bin", Float64.(getfield(data, f)))
    end
    writebin(stem * ".br.bin", Int64[b[q] for b in data.br for q in 1:2])      # row-major (nbr, 2)
    for (key, nm) in (("gen", "gen"), ("Sij", "sij"), ("Sji", "sji"), ("shunt", "sh"))
        ptr, idx = csr(data, key)
        writebin(stem * ".$(nm)_ptr.bin", ptr); writebin(stem * ".$(nm)_idx.bin", idx)
    end
    writebin(stem * ".lps.bin", Int64.(data.lps))
    writebin(stem * ".pgh_flat.bin", Float64[v for a in data.pgh for v in a]); writebin(stem * ".cgh_flat.bin", Float64[v for a in data.cgh for v in a])
    writebin(stem * ".pgh_len.bin", Int64[length(a) for a in data.pgh])
    writebin(stem * ".meta.bin", Int64[data.nbus, data.ngen, data.nbr, data.nbss, data.cost_order, data.slack])
    open(stem * ".source_type.txt", "w") do io
        print(io, data.source_type)
    end
end

for (cname, path) in CASES
    dump_data(cname, create_PowersenseData(path))
end

for (cname, path) in CASES, (fname, F) in FORMS
    stem = joinpath(DIR, "$(cname)_$(fname)")
    if !isfile(stem * ".x.bin")
        @warn "no evaluation point for $(cname) $(fname) (run `compare_julia_golden.py export` first)"
        continue
    end
    data = create_PowersenseData(path)
    # the flags of the committed fixtures (tests/golden/make_golden.py): run_opf!'s defaults with FACTS_bSeries off
    create_opf_model!(data, formulation = F, initialize = true, FACTS_bShunt = true, FACTS_bSeries = false,
                      box_constraints = false, obj_type = :linear)
    model = data.model
    d = JuMP.NLPEvaluator(model)
    MOI.initialize(d, [:Grad, :Jac, :Hess])                      # src/opt/MOI_wrapper.jl:1117
    n = JuMP.num_variables(model)
    m = JuMP.num_nl_constraints(model)
    x = readf64(stem * ".x.bin")
    mu = readf64(stem * ".mu.bin")
    length(x) == n || error("$(cname) $(fname): x has $(length(x)) entries, the model has $n variables")
    length(mu) == m || error("$(cname) $(fname): mu has $(length(mu)) entries, the model has $m NL rows")
    jstr = MOI.jacobian_structure(d)                             # :813
    hstr = MOI.hessian_lagrangian_structure(d)                   # :855
    g = zeros(m); J = zeros(length(jstr)); H = zeros(length(hstr))
    MOI.eval_constraint(d, g, x)                                 # :969
    MOI.eval_constraint_jacobian(d, J, x)                        # :1026
    MOI.eval_hessian_lagrangian(d, H, x, 1.0, mu)                # :1061 (no NL objective: sigma is unused)
    lo = Float64[c.lb for c in model.nlp_data.nlconstr]          # NLPBoundsPair of every row (:1091-1094)
    hi = Float64[c.ub for c in model.nlp_data.nlconstr]
    writebin(stem * ".g.bin", g); writebin(stem * ".J.bin", J); writebin(stem * ".H.bin", H)
    writebin(stem * ".lo.bin", lo); writebin(stem * ".hi.bin", hi)
    writebin(stem * ".jrow.bin", Int64[t[1] for t in jstr]); writebin(stem * ".jcol.bin", Int64[t[2] for t in jstr])
    writebin(stem * ".hrow.bin", Int64[t[1] for t in hstr]); writebin(stem * ".hcol.bin", Int64[t[2] for t in hstr])
    open(stem * ".meta.txt", "w") do io
        println(io, "n_var $n"); println(io, "m_nl $m"); println(io, "nnz_jac $(length(jstr))"); println(io, "nnz_hess $(length(hstr))")
        println(io, "julia $(VERSION)"); println(io, "JuMP $(pkgversion_of("JuMP"))")
        println(io, "MathOptInterface $(pkgversion_of("MathOptInterface"))"); println(io, "PowerModels $(pkgversion_of("PowerModels"))")
    end
    println("$(cname) $(fname): m_nl=$m nnzJ=$(length(jstr)) nnzH=$(length(hstr))")
end

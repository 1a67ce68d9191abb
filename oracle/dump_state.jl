# dump_state.jl — closes the "parity unpinned" gap on a machine that has Julia.
#
# Runs the REFERENCE itself (moschehaus/sph-mountain-waves, unmodified sources) on a small
# wcsph_perturbed_witch case and dumps what the north star wants bit-exact or within tolerance:
# cell keys, the neighbour pairs in traversal order, and x, v, rho, h after 0, 1 and NSTEPS steps.
# tests/test_julia_dump.py compares the oracle (and, on a GPU box, libsphmw) against the dump.
#
#   julia -t 4 oracle/dump_state.jl /path/to/sph-mountain-waves tests/golden/julia_dump 20 40e3 1000
#                                   reference checkout          output dir              n_y L   nsteps
#
# Not executable in the build image (no julia binary, SURVEY.md §8c); shipped for maintainers.
# The only edits made to the driver text before it is evaluated: `dr` and `dom_length` (the
# driver fixes them as module constants, wcsph_perturbed_witch.jl:26-27) and the plotting imports.

using Printf

ref_root = ARGS[1]
outdir = ARGS[2]
n_y = parse(Float64, ARGS[3])
dom_len = parse(Float64, ARGS[4])
nsteps = parse(Int, ARGS[5])
mkpath(outdir)

include(joinpath(ref_root, "src", "SmoothedParticles.jl"))
using .SmoothedParticles

src = read(joinpath(ref_root, "src", "current", "wcsph_perturbed_witch.jl"), String)
src = replace(src, "const dr = dom_height / 75" => "const dr = dom_height / $(n_y)")
src = replace(src, "const dom_length = 400e3" => "const dom_length = $(dom_len)")
src = replace(src, "using SmoothedParticles" => "using ..SmoothedParticles")
for dead in ("using DataFrames\n", "using Plots\n", "unicodeplots()\n",
             "include(joinpath(UTILS_DIR, \"new_packing.jl\"))\n")
    global src = replace(src, dead => "")
end
include_string(Main, src)
const W = Main.PerturbedStaticWitch

function write_f64(path, a)
    open(path, "w") do io
        write(io, Float64.(a))
    end
end
function write_i64(path, a)
    open(path, "w") do io
        write(io, Int64.(a))
    end
end

function dump(sys, tag)
    N = length(sys.particles)
    X = zeros(3, N); V = zeros(3, N)
    for (k, p) in enumerate(sys.particles), c in 1:3
        X[c, k] = p.x[c]; V[c, k] = p.v[c]
    end
    write_f64(joinpath(outdir, "x_$tag.f64"), X)       # 3 x N column-major = interleaved per particle
    write_f64(joinpath(outdir, "v_$tag.f64"), V)
    write_f64(joinpath(outdir, "rho_$tag.f64"), [p.ρ for p in sys.particles])
    write_f64(joinpath(outdir, "h_$tag.f64"), [p.h for p in sys.particles])
    write_f64(joinpath(outdir, "m_$tag.f64"), [p.m for p in sys.particles])
    write_f64(joinpath(outdir, "type_$tag.f64"), [p.type for p in sys.particles])
    write_i64(joinpath(outdir, "keys_$tag.i64"), [SmoothedParticles.find_key(sys, p.x) - 1 for p in sys.particles])
    # neighbour pairs in the traversal order of _apply_binary! (core.jl:94-112), 0-based
    index = IdDict(p => k for (k, p) in enumerate(sys.particles))
    pi = Int64[]; pj = Int64[]
    for p in sys.particles
        key = SmoothedParticles.find_key(sys, p.x)
        for dkey in sys.key_diff
            nk = key + dkey
            if 1 <= nk <= sys.key_max
                for j in sys.cell_list[nk].entries
                    j == 0 && break
                    q = sys.particles[j]
                    r = SmoothedParticles.dist(p, q)
                    ((r > sys.h) || (p === q)) && continue
                    push!(pi, index[p] - 1); push!(pj, j - 1)
                end
            end
        end
    end
    write_i64(joinpath(outdir, "pairs_i_$tag.i64"), pi)
    write_i64(joinpath(outdir, "pairs_j_$tag.i64"), pj)
    return N
end

sys = W.make_system()
counts = Dict{String,Int}()
counts["0"] = dump(sys, "0")
W.verlet_step!(sys)
counts["1"] = dump(sys, "1")
for k in 2:nsteps
    W.verlet_step!(sys)
end
counts[string(nsteps)] = dump(sys, string(nsteps))

open(joinpath(outdir, "meta.json"), "w") do io
    @printf(io, "{\"n_y\": %.17g, \"dom_length\": %.17g, \"nsteps\": %d, \"h\": %.17g, \"dt\": %.17g, ", n_y, dom_len, nsteps, sys.h, W.dt)
    @printf(io, "\"key_lim\": [%d, %d, %d], \"key_phase\": [%d, %d, %d], ", sys.key_lim..., sys.key_phase...)
    @printf(io, "\"n\": {%s}, \"threads\": %d, \"julia\": \"%s\"}\n",
            join(["\"$k\": $v" for (k, v) in counts], ", "), Threads.nthreads(), string(VERSION))
end
println("dumped to ", outdir)

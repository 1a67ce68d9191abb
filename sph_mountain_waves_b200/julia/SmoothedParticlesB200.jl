# SmoothedParticlesB200.jl — thin `ccall` shim over libsphmw.so (include/sphmw.h).
#
# Drop-in for the hot path of SmoothedParticles.jl as forked in moschehaus/sph-mountain-waves:
#   ParticleSystem / create_cell_list! / apply! / apply_unary! / apply_binary!   (src/structs.jl, src/core.jl)
#   wendland1/2/3, rDwendland*, spline*                                          (src/kernels.jl)
#   new_pvd_file / save_frame! / save_pvd_file                                   (src/IO.jl)
# A driver keeps its `mutable struct Particle <: AbstractParticle` and its closures; the closures
# are only used as *names* (Julia dispatches on the function object) and must belong to the
# device operator menu — there is no CPU fallback and no CUDA.jl compilation.
#
# NOT EXECUTED in the build image (no `julia` binary there); the Python ctypes binding
# (sph_mountain_waves_b200/_capi.py) exercises the very same symbols in tests/.
module SmoothedParticlesB200

using StaticArrays
export RealVector, VEC0, VECX, VECY, VECZ, AbstractParticle, ParticleSystem, register_operator!,
       create_cell_list!, apply!, apply_unary!, apply_binary!, verlet_steps!, sync_from_device!,
       new_pvd_file, save_frame!, save_pvd_file, wendland2, rDwendland2, wendland3, rDwendland3

const libsphmw = get(ENV, "LIBSPHMW", joinpath(@__DIR__, "..", "libsphmw.so"))

const RealVector = SVector{3,Float64}
const VEC0 = zero(RealVector); const VECX = RealVector(1,0,0); const VECY = RealVector(0,1,0); const VECZ = RealVector(0,0,1)
abstract type AbstractParticle end

struct SphmwConfig            # must match `struct sphmw_config`
    box_min::NTuple{3,Cdouble}
    box_max::NTuple{3,Cdouble}
    h::Cdouble
    capacity::Int64
    device::Int32
    flags::Int32
    slab_lo::Int64
    slab_hi::Int64
end

struct SphmwError <: Exception
    code::Int
    msg::String
end
function check(rc::Integer)
    rc < 0 && throw(SphmwError(rc, unsafe_string(ccall((:sphmw_last_error, libsphmw), Cstring, ()))))
    return rc
end

"""
    ParticleSystem(T, box_min, box_max, h; params, capacity, device=0, flags=0)

≙ ParticleSystem(T, domain, h) (structs.jl:57-91); pass `boundarybox(domain)` corners.
`params` are the driver's constants under libsphmw's names (dt, g, c, gamma, alpha, ...).
`particles` stays the host-side `Vector{T}`; device state is authoritative after `upload!`.
"""
mutable struct ParticleSystem{T<:AbstractParticle}
    h::Float64
    ctx::Ptr{Cvoid}
    particles::Vector{T}
    uploaded::Bool
    function ParticleSystem(T::DataType, box_min, box_max, h::Float64; params=Dict{String,Float64}(),
                            capacity::Int=0, device::Int=0, flags::Int=0)
        @assert(h > 0.0, "invalid ParticleSystem declaration! (h must be a positive float)")
        @assert(T <: AbstractParticle, "invalid ParticleSystem declaration! ("*string(T)*" is not an AbstractParticle subtype)")
        @assert(hasfield(T, :x) && fieldtype(T, :x) == RealVector, "invalid ParticleSystem declaration! (particles must have a field `x::RealVector`)")
        cfg = Ref(SphmwConfig(Tuple(box_min), Tuple(box_max), h, max(capacity, 1024), device, flags, -1, -1))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:sphmw_create, libsphmw), Cint, (Ref{SphmwConfig}, Ref{Ptr{Cvoid}}), cfg, out))
        sys = new{T}(h, out[], T[], false)
        for (k, v) in params
            check(ccall((:sphmw_set_param, libsphmw), Cint, (Ptr{Cvoid}, Cstring, Cdouble), sys.ctx, k, v))
        end
        finalizer(s -> ccall((:sphmw_destroy, libsphmw), Cint, (Ptr{Cvoid},), s.ctx), sys)
        return sys
    end
end

# Julia field name -> libsphmw field name (it accepts the unicode names as they are)
fieldname_c(s::Symbol) = String(s)

"AoS (vector of heap objects, structs.jl:53) -> SoA staging buffers -> device"
function upload!(sys::ParticleSystem{T}) where T
    N = length(sys.particles)
    check(ccall((:sphmw_resize, libsphmw), Cint, (Ptr{Cvoid}, Int64), sys.ctx, N))
    for f in fieldnames(T)
        FT = fieldtype(T, f)
        if FT == Float64
            buf = Float64[getfield(p, f) for p in sys.particles]
            GC.@preserve buf check(ccall((:sphmw_upload, libsphmw), Cint, (Ptr{Cvoid}, Cstring, Ptr{Cdouble}, Int64, Int32),
                                         sys.ctx, fieldname_c(f), buf, N, 1))
        elseif FT == RealVector
            buf = Matrix{Float64}(undef, N, 3)            # component-major: buf[i, c]
            for (i, p) in enumerate(sys.particles), c in 1:3
                buf[i, c] = getfield(p, f)[c]
            end
            GC.@preserve buf check(ccall((:sphmw_upload, libsphmw), Cint, (Ptr{Cvoid}, Cstring, Ptr{Cdouble}, Int64, Int32),
                                         sys.ctx, fieldname_c(f), buf, N, 3))
        end
    end
    sys.uploaded = true
end

"device -> host particles (call before reading `sys.particles`, e.g. at frame output)"
function sync_from_device!(sys::ParticleSystem{T}) where T
    n = Ref{Int64}(0)
    check(ccall((:sphmw_count, libsphmw), Cint, (Ptr{Cvoid}, Ref{Int64}), sys.ctx, n))
    N = n[]
    resize!(sys.particles, N)   # removal keeps the reference's swap-from-end order (core.jl:72-81)
    for f in fieldnames(T)
        FT = fieldtype(T, f)
        if FT == Float64
            buf = Vector{Float64}(undef, N)
            GC.@preserve buf check(ccall((:sphmw_download, libsphmw), Cint, (Ptr{Cvoid}, Cstring, Ptr{Cdouble}, Int64, Int32),
                                         sys.ctx, fieldname_c(f), buf, N, 1))
            for i in 1:N; setfield!(sys.particles[i], f, buf[i]); end
        elseif FT == RealVector
            buf = Matrix{Float64}(undef, N, 3)
            GC.@preserve buf check(ccall((:sphmw_download, libsphmw), Cint, (Ptr{Cvoid}, Cstring, Ptr{Cdouble}, Int64, Int32),
                                         sys.ctx, fieldname_c(f), buf, N, 3))
            for i in 1:N; setfield!(sys.particles[i], f, RealVector(buf[i,1], buf[i,2], buf[i,3])); end
        end
    end
end

"≙ create_cell_list!(sys) — core.jl:51-90"
function create_cell_list!(sys::ParticleSystem)
    sys.uploaded || upload!(sys)
    n = Ref{Int64}(0)
    check(ccall((:sphmw_create_cell_list, libsphmw), Cint, (Ptr{Cvoid}, Ref{Int64}), sys.ctx, n))
    return nothing
end

# closure -> operator name.  A driver registers its closures once:
#   register_operator!(compute_density!, "wcsph.compute_density")
const OPERATORS = IdDict{Function,String}()
register_operator!(f::Function, name::String) = (OPERATORS[f] = name; f)

"≙ apply!(sys, action!; self) — core.jl:151-161 (arity is a property of the device operator)"
function apply!(sys::ParticleSystem, action!::Function; self::Bool=false)
    haskey(OPERATORS, action!) || throw(SphmwError(-3, "closure $(action!) is not in the device operator menu (no CPU fallback)"))
    sys.uploaded || upload!(sys)
    check(ccall((:sphmw_apply, libsphmw), Cint, (Ptr{Cvoid}, Cstring, Int32), sys.ctx, OPERATORS[action!], self ? 1 : 0))
    return nothing
end
apply_unary!(sys::ParticleSystem, f::Function) = apply!(sys, f)
apply_binary!(sys::ParticleSystem, f::Function) = apply!(sys, f)

"fused fast path ≙ `for k in 1:n verlet_step!(sys) end` (wcsph_perturbed_witch.jl:309-332)"
function verlet_steps!(sys::ParticleSystem, scheme::String, n::Integer)
    sys.uploaded || upload!(sys)
    check(ccall((:sphmw_step, libsphmw), Cint, (Ptr{Cvoid}, Cstring, Int32), sys.ctx, scheme, n))
end

# flags of `ParticleSystem(...; flags)` (include/sphmw.h): the pair list records each particle's
# candidates on the first binary pass of a cell list and replays them on the later ones —
# apply!(sys, compute_density!) then apply!(sys, balance_of_momentum!) walk the cells once
const FLAG_FAST_MATH, FLAG_NO_PAIR_LIST, FLAG_PAIR_LIST_EAGER, FLAG_NO_PRETEST, FLAG_PACKED_RECORDS = 1, 4, 8, 16, 32
"(stride, lists built, particles that overflowed the stride, device bytes)"
function pair_list_info(sys::ParticleSystem)
    out = zeros(Int64, 4)
    GC.@preserve out check(ccall((:sphmw_pair_list_info, libsphmw), Cint, (Ptr{Cvoid}, Ptr{Int64}), sys.ctx, out))
    return Tuple(out)
end

# kernels.jl — evaluated on the device (scalar convenience wrappers)
function kernel_eval(name::String, h::Float64, r::Float64)
    hh = [h]; rr = [r]; out = [0.0]
    GC.@preserve hh rr out check(ccall((:sphmw_kernel_eval, libsphmw), Cint,
        (Cstring, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Int32), name, hh, rr, out, 1, 0))
    return out[1]
end
wendland2(h, r) = kernel_eval("wendland2", h, r);  rDwendland2(h, r) = kernel_eval("rDwendland2", h, r)
wendland3(h, r) = kernel_eval("wendland3", h, r);  rDwendland3(h, r) = kernel_eval("rDwendland3", h, r)

# IO.jl:20-75
struct DataStorage; sys::ParticleSystem; end
function new_pvd_file(sys::ParticleSystem, path::String)
    check(ccall((:sphmw_pvd_open, libsphmw), Cint, (Ptr{Cvoid}, Cstring), sys.ctx, path))
    return DataStorage(sys)
end
function save_frame!(data::DataStorage, sys::ParticleSystem, vars::Symbol...)
    names = [String(v) for v in vars]
    ptrs = [Base.unsafe_convert(Cstring, n) for n in names]
    GC.@preserve names check(ccall((:sphmw_pvd_save_frame, libsphmw), Cint, (Ptr{Cvoid}, Ptr{Cstring}, Int32),
                                   sys.ctx, ptrs, length(ptrs)))
end
save_pvd_file(data::DataStorage) = check(ccall((:sphmw_pvd_close, libsphmw), Cint, (Ptr{Cvoid},), data.sys.ctx))

end # module

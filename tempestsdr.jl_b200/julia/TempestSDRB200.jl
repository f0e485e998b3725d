# TempestSDRB200.jl -- ccall binding of libtempest_b200.so for TempestSDR.jl hosts.
#
# Julia is not installed in the build image, so this file has never been executed there; it is
# kept mechanical and mirrors, call for call, the ctypes binding in ../api.py that the GPU
# parity tests exercise (tests/test_abi.py checks every ccall's symbol, arity and argument types
# against include/tempest_b200.h).  Usage from the reference (see INTEGRATION.md):
#
#     include("TempestSDRB200.jl"); using .TempestSDRB200
#     TempestSDRB200.use!(TempestSDR)      # adds Float32 methods of amDemod, sig_to_image, ... that run on the GPU
#
# Every function keeps the reference signature (src/TempestSDR.jl:21-47 exports).  Matrices are
# plain column-major Julia Arrays, ComplexF32 vectors are passed as they are (interleaved re, im).
module TempestSDRB200

export amDemod, invert_amDemod, sig_to_image, downgradeImage, naiveResampler, init_resampler,
       calculate_autocorrelation, zoom_autocorr, SyncXY, vsync, Chain, push!, image, offsets,
       AtomicCircularBuffer, circ_put!, circ_take!, push_ring!, getSpectrum, getWelch, getWaterfall,
       Comm, comm_unique_id, allreduce!, integrate_device!

const LIB = get(ENV, "TEMPEST_B200_LIB", joinpath(@__DIR__, "..", "libtempest_b200.so"))
const RENDERING_SIZE = (600, 800)                      # src/GUI.jl:10

struct TempestB200Error <: Exception
    status::Cint
    msg::String
end
Base.showerror(io::IO, e::TempestB200Error) = print(io, "libtempest_b200 status $(e.status): $(e.msg)")

# Non-zero status -> exception (ErrorException-like), so the reference's try/catch blocks
# (src/GUI.jl:197-200) behave as before.  -5 is the reference's BoundsError.
function check(status::Cint)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:tsdr_last_error_string, LIB), Cstring, ()))
    status == -5 && throw(BoundsError())
    throw(TempestB200Error(status, msg))
end

# ---- Demodulation.jl -----------------------------------------------------------------------
function amDemod(sig::Array{ComplexF32})                                  # src/Demodulation.jl:26-28
    out = Array{Float32}(undef, size(sig))
    GC.@preserve sig out check(ccall((:tsdr_am_demod_f32, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t),
                                     pointer(sig), pointer(out), length(sig)))
    return out
end

function invert_amDemod(sig::Array{ComplexF32})                           # src/Demodulation.jl:31-35
    out = Array{Float32}(undef, size(sig))
    GC.@preserve sig out check(ccall((:tsdr_invert_am_demod_f32, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t),
                                     pointer(sig), pointer(out), length(sig)))
    return out
end

# ---- Resampler.jl --------------------------------------------------------------------------
function sig_to_image(sig::AbstractVector{Float32}, y_t, x_t)              # src/Resampler.jl:117-122
    s = sig isa Vector{Float32} ? sig : collect(sig)                       # views are copied once
    out = Matrix{Float32}(undef, Int(y_t), Int(x_t))
    GC.@preserve s out check(ccall((:tsdr_sig_to_image_f32, LIB), Cint,
                                   (Ptr{Cvoid}, Csize_t, Cint, Cint, Ptr{Cvoid}),
                                   pointer(s), length(s), y_t, x_t, pointer(out)))
    return out
end

function downgradeImage(image::Matrix{Float32})                            # src/Resampler.jl:124-126
    out = Matrix{Float32}(undef, RENDERING_SIZE...)
    GC.@preserve image out check(ccall((:tsdr_downgrade_f32, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}),
                                       pointer(image), size(image, 1), size(image, 2), pointer(out)))
    return out
end

function naiveResampler(sigOut::Vector{Float32}, sigId::Vector{Float32}, upCoeff)   # src/Resampler.jl:103-110
    length(sigOut) >= upCoeff * length(sigId) || throw(BoundsError(sigOut, upCoeff * length(sigId)))
    GC.@preserve sigOut sigId check(ccall((:tsdr_naive_resampler_f32, LIB), Cint,
                                          (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Cint),
                                          pointer(sigOut), pointer(sigId), length(sigId), upCoeff))
    return nothing
end

# init_resampler(T, bufferSize, upCoeff) -> resampler!(out, in)                 src/Resampler.jl:26-62
function init_resampler(T::Type, bufferSize::Int, upCoeff::Int)
    T === Float32 || error("libtempest_b200 implements init_resampler for Float32 only")
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:tsdr_upsampler_create, LIB), Cint, (Csize_t, Cint, Ptr{Ptr{Cvoid}}), bufferSize, upCoeff, h))
    handle = h[]
    function resampler!(out::AbstractVector{T2}, in::AbstractVector{T2}) where T2
        @assert T == T2 "Type of input ($T2) should match type used during init ($T)"
        @assert length(in) == bufferSize "Size of input $(length(in)) should match size used during init $bufferSize"
        GC.@preserve out in check(ccall((:tsdr_upsampler_apply_f32, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}, Csize_t),
                                        handle, pointer(out), length(out), pointer(in), length(in)))
    end
    return resampler!
end
init_resampler(x::Vector{T}, upCoeff) where T = init_resampler(T, length(x), upCoeff)

# ---- Autocorrelations.jl -------------------------------------------------------------------
function calculate_autocorrelation(x::Vector{Float32}, Fs, minDelay, maxDelay, scale = :log)   # :23-37
    n = Ref{Csize_t}(0)
    check(ccall((:tsdr_autocorr_out_len, LIB), Cint, (Csize_t, Cdouble, Cdouble, Cdouble, Ptr{Csize_t}),
                length(x), Fs, minDelay, maxDelay, n))
    out = Vector{Float32}(undef, n[])
    GC.@preserve x out check(ccall((:tsdr_autocorr_f32, LIB), Cint,
                                   (Ptr{Cvoid}, Csize_t, Cdouble, Cdouble, Cdouble, Cint, Ptr{Cvoid}, Ptr{Csize_t}),
                                   pointer(x), length(x), Fs, minDelay, maxDelay, scale == :log ? 1 : 0, pointer(out), n))
    indexMin = 1 + Int(round(minDelay * Fs)); indexMax = Int(round(maxDelay * Fs))
    lags = (0:(indexMax - indexMin)) * 1 / Fs
    return out, lags
end

# zoom_autocorr is index arithmetic on the host; it stays as in src/Autocorrelations.jl:42-53.
function zoom_autocorr(Γ, Fs; rate_min = 20, rate_max = 100)
    N = length(Γ)
    pos_rate_min = min(Int(round(1 / rate_max * Fs)), N)
    pos_rate_max = min(Int(round(1 / rate_min * Fs)), N)
    return 1 ./ ((pos_rate_min:pos_rate_max) ./ Fs), Γ[pos_rate_min:pos_rate_max]
end

# ---- FrameSynchronisation.jl ---------------------------------------------------------------
mutable struct SyncXY{T}                                                   # :25-48 ; state lives on the GPU
    handle::Ptr{Cvoid}
    n_y::Int
    n_x::Int
    function SyncXY(image::Matrix{T}) where T
        T === Float32 || error("libtempest_b200 implements SyncXY for Float32 images")
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:tsdr_sync_create, LIB), Cint, (Cint, Cint, Ptr{Ptr{Cvoid}}), size(image, 1), size(image, 2), h))
        s = new{T}(h[], size(image, 1), size(image, 2))
        finalizer(x -> ccall((:tsdr_sync_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), s)
        return s
    end
end

function vsync(image::AbstractMatrix{Float32}, sync::SyncXY{Float32})      # :56-79 -> (s_y, s_x), 1-based
    img = image isa Matrix{Float32} ? image : collect(image)
    sy = Ref{Cint}(0); sx = Ref{Cint}(0)
    GC.@preserve img check(ccall((:tsdr_vsync_f32, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cint}, Ptr{Cint}),
                                 sync.handle, pointer(img), sy, sx))
    return (Int(sy[]), Int(sx[]))
end

# β_x / β_y of the Julia struct, copied out on demand (column-major, (1+wmax-wmin) x n)
function betas(sync::SyncXY{Float32})
    b1 = Ref{Cint}(0); b2 = Ref{Cint}(0); b3 = Ref{Cint}(0); b4 = Ref{Cint}(0)   # ccall takes no splatted arguments
    check(ccall((:tsdr_sync_bounds, LIB), Cint, (Ptr{Cvoid}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}), sync.handle, b1, b2, b3, b4))
    wmin_y, wmax_y, wmin_x, wmax_x = Int(b1[]), Int(b2[]), Int(b3[]), Int(b4[])
    βx = Matrix{Float32}(undef, 1 + wmax_x - wmin_x, sync.n_x); βy = Matrix{Float32}(undef, 1 + wmax_y - wmin_y, sync.n_y)
    GC.@preserve βx βy check(ccall((:tsdr_sync_get_beta, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                                   sync.handle, pointer(βx), pointer(βy)))
    return βx, βy
end

# ---- the fused loop body of coreProcessing (src/GUI.jl:163-178) ------------------------------
mutable struct Chain
    handle::Ptr{Cvoid}
    function Chain(Fs, x_t, y_t, fv; alpha = 0.1f0, max_samples, device = 0, flags = 0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:tsdr_chain_create, LIB), Cint,
                    (Ptr{Ptr{Cvoid}}, Cint, Cdouble, Cint, Cint, Cdouble, Cfloat, Csize_t, Cuint, Ptr{Cvoid}),
                    h, device, Fs, x_t, y_t, fv, alpha, max_samples, flags, C_NULL))
        c = new(h[])
        finalizer(x -> ccall((:tsdr_chain_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), c)
        return c
    end
end

# one recv! buffer: amDemod -> sig_to_image -> downgradeImage -> vsync -> circshift -> EMA for every frame
# a buffer read from a `:short` recording without the host-side widening of readComplexBinary
# (src/DatBinaryFiles.jl:47-49): raw = reinterpret(Int16, read(file)) holds (re, im) pairs
function push_int16!(c::Chain, raw::Vector{Int16})
    iseven(length(raw)) || throw(ArgumentError("Int16 IQ buffer needs (re, im) pairs"))
    n = Ref{Cint}(0)
    GC.@preserve raw begin
        check(ccall((:tsdr_chain_push_host_i16, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cint}),
                    c.handle, pointer(raw), length(raw) ÷ 2, n))
        check(ccall((:tsdr_chain_sync, LIB), Cint, (Ptr{Cvoid},), c.handle))
    end
    return Int(n[])
end

function Base.push!(c::Chain, sigId::Vector{ComplexF32})
    n = Ref{Cint}(0)
    GC.@preserve sigId begin
        check(ccall((:tsdr_chain_push_host, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cint}),
                    c.handle, pointer(sigId), length(sigId), n))
        check(ccall((:tsdr_chain_sync, LIB), Cint, (Ptr{Cvoid},), c.handle))   # sigId may be reused after return
    end
    return Int(n[])
end

function image(c::Chain)                                                   # imageOut, 600 x 800
    out = Matrix{Float32}(undef, RENDERING_SIZE...)
    GC.@preserve out check(ccall((:tsdr_chain_read_image, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), c.handle, pointer(out)))
    return out
end

function offsets(c::Chain, maxframes = 4096)
    sy = Vector{Cint}(undef, maxframes); sx = Vector{Cint}(undef, maxframes); n = Ref{Cint}(0)
    GC.@preserve sy sx check(ccall((:tsdr_chain_read_offsets, LIB), Cint, (Ptr{Cvoid}, Ptr{Cint}, Ptr{Cint}, Cint, Ptr{Cint}),
                                   c.handle, pointer(sy), pointer(sx), maxframes, n))
    k = min(Int(n[]), maxframes)
    return Int.(sy[1:k]), Int.(sx[1:k])
end

# per frame of the last buffer: (max beta_x, max beta_y, Sigma_x, Sigma_y) -- what a configuration search scores
function scores(c::Chain, maxframes = 4096)
    bx = zeros(Float32, maxframes); by = zeros(Float32, maxframes)
    sx = zeros(Float32, maxframes); sy = zeros(Float32, maxframes)
    n = Ref{Cint}(0)
    GC.@preserve bx by sx sy check(ccall((:tsdr_chain_read_scores, LIB), Cint,
                                         (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cint}),
                                         c.handle, pointer(bx), pointer(by), pointer(sx), pointer(sy), maxframes, n))
    k = min(Int(n[]), maxframes)
    return bx[1:k], by[1:k], sx[1:k], sy[1:k]
end

configure!(c::Chain, Fs, x_t, y_t, fv) = check(ccall((:tsdr_chain_configure, LIB), Cint,
                                                     (Ptr{Cvoid}, Cdouble, Cint, Cint, Cdouble), c.handle, Fs, x_t, y_t, fv))
set_alpha!(c::Chain, α) = check(ccall((:tsdr_chain_set_alpha, LIB), Cint, (Ptr{Cvoid}, Cfloat), c.handle, α))

# ---- GetSpectrum.jl (src/GetSpectrum.jl:21-66) -------------------------------------------------------------
function getSpectrum(fs, sig::Vector{ComplexF32}; N = nothing)
    isnothing(N) && (N = length(sig))
    N <= length(sig) || throw(BoundsError(sig, 1:N))
    freqAx = collect(((0:N-1) ./ N .- 0.5) * fs)
    y = Vector{Float32}(undef, N)
    GC.@preserve sig y check(ccall((:tsdr_get_spectrum_f32, LIB), Cint, (Ptr{Cvoid}, Csize_t, Cint, Ptr{Cvoid}),
                                   pointer(sig), N, 1, pointer(y)))
    return (freqAx, y)
end
getSpectrum(sig) = getSpectrum(1, sig)

function getWelch(fe, sig::Vector{ComplexF32}; sizeFFT = 1024)
    y = Vector{Float32}(undef, sizeFFT)
    GC.@preserve sig y check(ccall((:tsdr_get_welch_f32, LIB), Cint, (Ptr{Cvoid}, Csize_t, Cint, Ptr{Cvoid}),
                                   pointer(sig), length(sig), sizeFFT, pointer(y)))
    return (collect(((0:sizeFFT-1) ./ sizeFFT .- 0.5) * fe), y)
end

function getWaterfall(fe, sig::Vector{ComplexF32}; sizeFFT = 1024)
    nbSeg = length(sig) ÷ sizeFFT
    s = Matrix{Float32}(undef, sizeFFT, nbSeg)
    GC.@preserve sig s check(ccall((:tsdr_get_waterfall_f32, LIB), Cint, (Ptr{Cvoid}, Csize_t, Cint, Ptr{Cvoid}),
                                   pointer(sig), length(sig), sizeFFT, pointer(s)))
    fAx = collect(((0:1:sizeFFT-1) ./ sizeFFT .- 0.5) .* fe)
    tAx = (0:nbSeg-1) * (sizeFFT / fe)
    return tAx, fAx, Float64.(s)
end
getWaterfall(sig; sizeFFT = 1024) = getWaterfall(1, sig; sizeFFT = sizeFFT)

# ---- AtomicCircularBuffer (src/AtomicAbstractSDRs.jl:67-190) in page-locked memory ------------------------
# Same constructor and circ_put! / circ_take! as the reference, so start_atomic_sdr (:284-306) and recv!
# (:312-314) keep working unchanged; push_ring!(chain, ring) replaces recv! + the loop body (GUI.jl:163-176)
# and copies to the GPU straight from the ring's slot.
mutable struct AtomicCircularBuffer{T}
    handle::Ptr{Cvoid}
    nEch::Int
    depth::Int
    function AtomicCircularBuffer{T}(nEch::Int, depth::Int) where T
        T === ComplexF32 || T === Complex{Int16} || throw(ArgumentError("ring slots hold ComplexF32 or Complex{Int16}"))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:tsdr_ring_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Csize_t, Cint, Cint), h, nEch * sizeof(T), depth, 1))
        r = new{T}(h[], nEch, depth)
        finalizer(x -> ccall((:tsdr_ring_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), r)
        return r
    end
end

function circ_put!(circ_buff::AtomicCircularBuffer{T}, data::Vector{T}) where T          # :159-170
    GC.@preserve data check(ccall((:tsdr_ring_put, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t),
                                  circ_buff.handle, pointer(data), sizeof(data)))
end

function circ_take!(buffer::Vector{T}, circ_buff::AtomicCircularBuffer{T}) where T       # :176-189
    # @threadcall: the wait for new data must not block the Julia scheduler the producer task runs on
    GC.@preserve buffer check(@threadcall((:tsdr_ring_take, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Cint),
                                          circ_buff.handle, pointer(buffer), sizeof(buffer), -1))
end

function push_ring!(c::Chain, ring::AtomicCircularBuffer{T}; timeout_ms = -1) where T
    n = Ref{Cint}(0)
    fmt = T === ComplexF32 ? 0 : 1
    check(@threadcall((:tsdr_chain_push_ring, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Cint}),
                      c.handle, ring.handle, fmt, timeout_ms, n))
    return Int(n[])
end

# ---- multi-GPU combine of a long integration (BASELINE cfg 5; src/GUI.jl:175 is linear in the frames) --------------
# One communicator per GPU.  One Julia process per GPU (Distributed / MPI.jl / several julia's): rank 0 calls
# comm_unique_id(), ships the 128 bytes to the others by any means, every rank builds Comm(id, nranks, rank; device).
comm_unique_id() = (id = Vector{UInt8}(undef, 128);
                    GC.@preserve id check(ccall((:tsdr_comm_get_unique_id, LIB), Cint, (Ptr{Cvoid},), pointer(id))); id)

mutable struct Comm
    handle::Ptr{Cvoid}
    nranks::Int
    rank::Int
    function Comm(id::Vector{UInt8}, nranks::Integer, rank::Integer; device = 0)
        length(id) == 128 || throw(ArgumentError("unique id must be 128 bytes"))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve id check(ccall((:tsdr_comm_init_rank, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint, Cint, Cint, Ptr{Cvoid}),
                                    h, device, nranks, rank, pointer(id)))
        c = new(h[], nranks, rank)
        finalizer(x -> ccall((:tsdr_comm_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), c)
        return c
    end
    function Comm(handle::Ptr{Cvoid}, nranks::Integer, rank::Integer)      # a communicator made by comm_init_all
        c = new(handle, nranks, rank)
        finalizer(x -> ccall((:tsdr_comm_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), c)
        return c
    end
end

# ONE Julia process driving several GPUs of a box: a communicator per device (devices 0..n-1, or the ones listed);
# collectives issued from one thread for several of them go between group_start() and group_end()
function comm_init_all(n::Integer; devices::Union{Vector{Cint},Nothing} = nothing)
    hs = Vector{Ptr{Cvoid}}(undef, n)
    devices === nothing || length(devices) == n || throw(ArgumentError("devices must list n entries"))
    GC.@preserve hs devices check(ccall((:tsdr_comm_init_all, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint, Ptr{Cint}),
                                        pointer(hs), n, devices === nothing ? Ptr{Cint}(C_NULL) : pointer(devices)))
    return [Comm(hs[r], n, r - 1) for r in 1:n]
end
group_start() = check(ccall((:tsdr_comm_group_start, LIB), Cint, ()))
group_end() = check(ccall((:tsdr_comm_group_end, LIB), Cint, ()))

# imageOut <- sum over ranks of weight_rank * imageOut_rank (weight = α^(frames after this rank's block)); asynchronous
allreduce!(c::Chain, comm::Comm, weight = 1.0f0) =
    check(ccall((:tsdr_chain_allreduce, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cfloat), c.handle, comm.handle, weight))

# one rank's share of a sharded integration in one call: reset, prime with the halo frame, push the device buffers
# (CuArray pointers or any device addresses), combine.  halo == C_NULL for the first block, comm === nothing for 1 GPU.
function integrate_device!(c::Chain, halo::Ptr{Cvoid}, halo_samples::Integer, bufs::Vector{Ptr{Cvoid}}, samples::Vector{Csize_t};
                           comm::Union{Comm,Nothing} = nothing, weight = 1.0f0)
    n = Ref{Cint}(0)
    GC.@preserve bufs samples check(ccall((:tsdr_chain_integrate_device, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Ptr{Cvoid}}, Ptr{Csize_t}, Cint, Ptr{Cvoid}, Cfloat, Ptr{Cint}),
        c.handle, halo, halo_samples, pointer(bufs), pointer(samples), length(bufs),
        comm === nothing ? C_NULL : comm.handle, weight, n))
    return Int(n[])
end

# ---- vsync on the REFERENCE's own SyncXY{Float32} -------------------------------------------------------------------
# coreProcessing builds the reference's struct itself (src/GUI.jl:136) and hands it to vsync (:171).  Its whole state is
# the pair of public tables β_x / β_y (src/FrameSynchronisation.jl:25-30): s_y is read from the β_y the PREVIOUS call left
# there (:66), so that part stays on the host exactly as the reference does it, the GPU computes both new tables and
# s_x, and the tables are copied back into the struct's arrays.  Device handles are stateless here: one per image size.
const _SYNC_POOL = Dict{Tuple{Int,Int},Tuple{Ptr{Cvoid},ReentrantLock}}()
const _SYNC_POOL_LOCK = ReentrantLock()
function _pooled_sync(n_y::Int, n_x::Int)
    lock(_SYNC_POOL_LOCK) do
        get!(_SYNC_POOL, (n_y, n_x)) do
            h = Ref{Ptr{Cvoid}}(C_NULL)
            check(ccall((:tsdr_sync_create, LIB), Cint, (Cint, Cint, Ptr{Ptr{Cvoid}}), n_y, n_x, h))
            (h[], ReentrantLock())
        end
    end
end

function vsync_into(image::AbstractMatrix{Float32}, sync)          # sync: anything with Float32 matrices β_x, β_y
    img = image isa Matrix{Float32} ? image : collect(image)
    s_y = findmax(sync.β_y)[2][2]                                  # :66 -- the table of the previous call
    (handle, lk) = _pooled_sync(size(img, 1), size(img, 2))
    sy = Ref{Cint}(0); sx = Ref{Cint}(0)
    βx = sync.β_x; βy = sync.β_y
    lock(lk) do
        GC.@preserve img βx βy begin
            check(ccall((:tsdr_vsync_f32, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cint}, Ptr{Cint}), handle, pointer(img), sy, sx))
            check(ccall((:tsdr_sync_get_beta, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), handle, pointer(βx), pointer(βy)))
        end
    end
    return (s_y, Int(sx[]))
end

# ---- drop-in: zero changes to GUI.jl ---------------------------------------------------------------------------------
# The reference's DSP functions are untyped or generic in T (src/Demodulation.jl:26, src/Resampler.jl:117,124,
# src/Autocorrelations.jl:23, src/FrameSynchronisation.jl:56).  use! adds, IN THE MODULE THAT OWNS EACH FUNCTION, a method
# for the Float32 argument types the GUI actually passes (GUI.jl:128-129,164-171; :64-73).  Those methods are strictly
# more specific than the reference's, so dispatch picks them for Float32 data and every other element type keeps the
# reference's own code -- nothing is overwritten, no name is rebound.  (The functions reach TempestSDR through
# `@reexport using .Resampler` etc., src/TempestSDR.jl:27-46, which is why the methods must be evaluated in the submodules.)
function use!(ref::Module)
    B = @__MODULE__
    Core.eval(ref, quote                                   # Demodulation.jl is included at TempestSDR's top level (:21-23)
        amDemod(sig::Array{ComplexF32}) = $B.amDemod(sig)
        invert_amDemod(sig::Array{ComplexF32}) = $B.invert_amDemod(sig)
    end)
    Core.eval(ref.Resampler, quote
        sig_to_image(sig::AbstractVector{Float32}, y_t, x_t) = $B.sig_to_image(sig, y_t, x_t)
        downgradeImage(image::Matrix{Float32}) = $B.downgradeImage(image)
        naiveResampler(sigOut::Vector{Float32}, sigId::Vector{Float32}, upCoeff) = $B.naiveResampler(sigOut, sigId, upCoeff)
        init_resampler(::Type{Float32}, bufferSize::Int, upCoeff::Int) = $B.init_resampler(Float32, bufferSize, upCoeff)
    end)
    Core.eval(ref.Autocorrelations, quote
        calculate_autocorrelation(x::Vector{Float32}, Fs, minDelay, maxDelay, scale = :log) =
            $B.calculate_autocorrelation(x, Fs, minDelay, maxDelay, scale)
    end)
    Core.eval(ref.FrameSynchronisation, quote                # the reference's own SyncXY{Float32}: tables stay in the struct
        vsync(image::AbstractMatrix{Float32}, sync::SyncXY{Float32}) = $B.vsync_into(image, sync)
    end)
    Core.eval(ref.GetSpectrum, quote
        getSpectrum(fs, sig::Vector{ComplexF32}; N = nothing) = $B.getSpectrum(fs, sig; N = N)
        getWelch(fe, sig::Vector{ComplexF32}; sizeFFT = 1024) = $B.getWelch(fe, sig; sizeFFT = sizeFFT)
        getWaterfall(fe, sig::Vector{ComplexF32}; sizeFFT = 1024) = $B.getWaterfall(fe, sig; sizeFFT = sizeFFT)
    end)
    return nothing
end

end # module

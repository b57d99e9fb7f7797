# CUDABackend.jl -- the Julia side of the drop-in: a module to `include` from Kinematics.jl
# (after collision.jl) that binds libkin_b200.so with `ccall` and adds batched methods to the
# reference's own exported functions.  The reference's single-configuration API is untouched.
#
# NOT EXECUTED in the build image (Julia is not installed there); it mirrors, call for call, the Python
# host mirror in kinematics.jl_b200/ (device.py, algorithm.py, collision.py), which IS tested on B200.
#
#   dm = CUDABackend.DeviceMechanism(mech, joints; sscc=sscc, sdf=sdf)
#   Q  = CUDA.rand(Float64, N, n_dof)                  # SoA: Julia (N, n_dof) column-major == q[c*ld + n]
#   T  = get_transform(dm, links, Q)                   # (N, 12, length(links))
#   J  = get_jacobian(dm, link, Q, true; rpy_jac=false)# (N, rows, n_dof)
#   v, g = compute_coll_dists_and_grads(dm, Q; truncation_dist=Inf)   # (N, S), (N, n_dof, S)
module CUDABackend

using CUDA
using ..Kinematics: Mechanism, Link, Joint, Fixed, Revolute, Prismatic, Transform, SweptSphereCollisionChecker,
                    AbstractSDF, BoxSDF, UnionSDF, parent_joint, isroot, get_transform
import ..Kinematics: get_transform, get_jacobian, compute_coll_dists, compute_coll_dists_and_grads

const libkin = get(ENV, "KIN_B200_LIB", "libkin_b200.so")

const KIN_F64, KIN_F32 = Cint(0), Cint(1)
const KIN_LAYOUT_SOA, KIN_LAYOUT_AOS, KIN_LAYOUT_TILED32 = Cint(0), Cint(1), Cint(2)
const KIN_GRAD_FD, KIN_GRAD_ANALYTIC, KIN_GRAD_FD_DIRECT = Cint(0), Cint(1), Cint(2)
const KIN_SCRATCH_REFERENCE, KIN_SCRATCH_CLEAN = Cint(0), Cint(1)

# include/kin_b200.h: KinModelDesc
struct KinModelDesc
    n_links::Cint
    parent_link::Ptr{Cint}
    joint_type::Ptr{Cint}
    joint_pose::Ptr{Cdouble}
    joint_axis::Ptr{Cdouble}
    q_index::Ptr{Cint}
    default_angle::Ptr{Cdouble}
    n_joints::Cint
    with_base::Cint
    n_spheres::Cint
    sphere_link::Ptr{Cint}
    sphere_center::Ptr{Cdouble}
    sphere_radius::Ptr{Cdouble}
    n_boxes::Cint
    box_pose::Ptr{Cdouble}
    box_width::Ptr{Cdouble}
end

# include/kin_b200.h: KinCall
struct KinCall
    precision::Cint
    layout::Cint
    n::Int64
    batch_stride::Int64
    q::CuPtr{Cvoid}
    n_fk_links::Cint
    fk_links::Ptr{Cint}
    T_out::CuPtr{Cvoid}
    n_jac_links::Cint
    jac_links::Ptr{Cint}
    with_rot::Cint
    rpy_jac::Cint
    keep_irrelevant::Cint
    J_out::CuPtr{Cvoid}
    truncation_dist::Cdouble
    grad_mode::Cint
    scratch_mode::Cint
    vals_out::CuPtr{Cvoid}
    grads_out::CuPtr{Cvoid}
    argmin_out::CuPtr{Cint}
    vals_offset::Cdouble
    stream::Ptr{Cvoid}
end

check(rc) = rc == 0 || error("libkin_b200: " * unsafe_string(ccall((:kin_last_error, libkin), Cstring, ())))

joint_type_code(::Joint{Fixed}) = Cint(0)
joint_type_code(::Joint{Revolute}) = Cint(1)
joint_type_code(::Joint{Prismatic}) = Cint(2)
joint_axis(j::Joint{Fixed}) = (0.0, 0.0, 0.0)
joint_axis(j::Joint) = Tuple(j.jt.axis)

mutable struct DeviceMechanism
    handle::Ptr{Cvoid}
    n_dof::Int
    n_spheres::Int
    function DeviceMechanism(m::Mechanism, joints::Vector{<:Joint};
                             sscc::Union{SweptSphereCollisionChecker, Nothing}=nothing,
                             sdf::Union{AbstractSDF, Nothing}=nothing)
        L = length(m.links)
        parent = fill(Cint(-1), L); jtype = zeros(Cint, L); qidx = fill(Cint(-1), L)
        pose = zeros(Cdouble, 16, L); axis = zeros(Cdouble, 3, L); defang = zeros(Cdouble, L)
        col = Dict(j.id => Cint(c - 1) for (c, j) in enumerate(joints))
        for l in m.links
            isroot(l) && continue
            j = parent_joint(m, l)
            parent[l.id] = l.plink_id
            jtype[l.id] = joint_type_code(j)
            pose[:, l.id] = vec(j.pose.mat)            # Transform.mat is already column-major 4x4
            axis[:, l.id] .= joint_axis(j)
            qidx[l.id] = get(col, j.id, Cint(-1))
            defang[l.id] = m.angles[j.id]
        end
        # spheres: sscc.sphere_links are children of the collision link through a pure translation
        S = sscc === nothing ? 0 : length(sscc.sphere_links)
        slink = Cint[sscc.sphere_links[i].plink_id for i in 1:S]
        scen = zeros(Cdouble, 3, S)
        for i in 1:S
            scen[:, i] = m.joints[sscc.sphere_links[i].pjoint_id].pose.mat[1:3, 4]
        end
        srad = S == 0 ? Cdouble[] : Vector{Cdouble}(sscc.sphere_radii)
        boxes = sdf === nothing ? AbstractSDF[] : (sdf isa UnionSDF ? sdf.sdfs : [sdf])
        B = length(boxes)
        bpose = zeros(Cdouble, 16, B); bwidth = zeros(Cdouble, 3, B)
        for (i, b) in enumerate(boxes)
            Kinematics.inv_pose(b)                      # refreshes b.pose for attached boxes (sdf.jl:24-32)
            bpose[:, i] = vec(b.pose.mat); bwidth[:, i] .= b.width
        end
        handle = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve parent jtype pose axis qidx defang slink scen srad bpose bwidth begin
            desc = KinModelDesc(L, pointer(parent), pointer(jtype), pointer(pose), pointer(axis), pointer(qidx),
                                pointer(defang), length(joints), m.with_base, S, pointer(slink), pointer(scen),
                                pointer(srad), B, pointer(bpose), pointer(bwidth))
            check(ccall((:kin_model_create, libkin), Cint, (Ref{KinModelDesc}, Ref{Ptr{Cvoid}}), desc, handle))
        end
        dm = new(handle[], length(joints) + (m.with_base ? 3 : 0), S)
        finalizer(x -> ccall((:kin_model_destroy, libkin), Cint, (Ptr{Cvoid},), x.handle), dm)
        return dm
    end
end

function eval!(dm::DeviceMechanism, Q::CuMatrix{Float64}; fk_links=Cint[], T=nothing, jac_links=Cint[], J=nothing,
               with_rot=true, rpy_jac=false, vals=nothing, grads=nothing, argmin=nothing,
               truncation_dist=Inf, grad_mode=KIN_GRAD_FD, scratch_mode=KIN_SCRATCH_REFERENCE, vals_offset=0.0)
    N = size(Q, 1)
    @assert size(Q, 2) == dm.n_dof
    p(x) = x === nothing ? CU_NULL : pointer(x)
    GC.@preserve fk_links jac_links begin
        call = KinCall(KIN_F64, KIN_LAYOUT_SOA, N, 0, pointer(Q), length(fk_links), pointer(fk_links), p(T),
                       length(jac_links), pointer(jac_links), with_rot, rpy_jac, 0, p(J), truncation_dist, grad_mode,
                       scratch_mode, p(vals), p(grads), argmin === nothing ? CU_NULL : pointer(argmin), vals_offset,
                       Base.unsafe_convert(Ptr{Cvoid}, CUDA.stream().handle))
        check(ccall((:kin_eval, libkin), Cint, (Ptr{Cvoid}, Ref{KinCall}), dm.handle, call))
    end
end

# get_transform(m, link) for a batch: (N, 12, n_links), each transform 3x4 column-major (algorithm.jl:1)
function get_transform(dm::DeviceMechanism, links::Vector{<:Link}, Q::CuMatrix{Float64})
    ids = Cint[l.id for l in links]
    T = CuArray{Float64}(undef, size(Q, 1), 12, length(links))
    eval!(dm, Q; fk_links=ids, T=T)
    return T
end

# get_jacobian(m, link, joints, with_rot; rpy_jac) for a batch: (N, rows, n_dof) (algorithm.jl:108)
function get_jacobian(dm::DeviceMechanism, link::Link, Q::CuMatrix{Float64}, with_rot::Bool; rpy_jac=false)
    J = CuArray{Float64}(undef, size(Q, 1), with_rot ? 6 : 3, dm.n_dof)
    eval!(dm, Q; jac_links=Cint[link.id], J=J, with_rot=with_rot, rpy_jac=rpy_jac)
    return J
end

# compute_coll_dists (collision.jl:60): (N, S)
function compute_coll_dists(dm::DeviceMechanism, Q::CuMatrix{Float64})
    vals = CuArray{Float64}(undef, size(Q, 1), dm.n_spheres)
    eval!(dm, Q; vals=vals)
    return vals
end

# compute_coll_dists_and_grads (collision.jl:96): (N, S), (N, n_dof, S)
function compute_coll_dists_and_grads(dm::DeviceMechanism, Q::CuMatrix{Float64}; truncation_dist=Inf,
                                      grad_mode=KIN_GRAD_FD, scratch_mode=KIN_SCRATCH_REFERENCE, margin=0.0)
    vals = CuArray{Float64}(undef, size(Q, 1), dm.n_spheres)
    grads = CuArray{Float64}(undef, size(Q, 1), dm.n_dof, dm.n_spheres)
    eval!(dm, Q; vals=vals, grads=grads, truncation_dist=truncation_dist, grad_mode=grad_mode,
          scratch_mode=scratch_mode, vals_offset=margin)
    return vals, grads
end

end # module

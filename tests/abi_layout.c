/* Prints the layout gcc gives the structs of include/kin_b200.h as JSON (tests/test_julia_binding.py compares it
 * with the struct declarations of julia/CUDABackend.jl and with the ctypes structures of kinematics.jl_b200/lib.py). */
#include <stddef.h>
#include <stdio.h>

#include "kin_b200.h"

#define F(S, f) printf("%s    {\"name\": \"%s\", \"offset\": %zu, \"size\": %zu}", first ? "" : ",\n", #f, offsetof(S, f), sizeof(((S *)0)->f)), first = 0

int main(void) {
    int first = 1;
    printf("{\n  \"KinModelDesc\": {\"size\": %zu, \"fields\": [\n", sizeof(KinModelDesc));
    F(KinModelDesc, n_links); F(KinModelDesc, parent_link); F(KinModelDesc, joint_type); F(KinModelDesc, joint_pose);
    F(KinModelDesc, joint_axis); F(KinModelDesc, q_index); F(KinModelDesc, default_angle); F(KinModelDesc, n_joints);
    F(KinModelDesc, with_base); F(KinModelDesc, n_spheres); F(KinModelDesc, sphere_link); F(KinModelDesc, sphere_center);
    F(KinModelDesc, sphere_radius); F(KinModelDesc, n_boxes); F(KinModelDesc, box_pose); F(KinModelDesc, box_width);
    first = 1;
    printf("\n  ]},\n  \"KinCall\": {\"size\": %zu, \"fields\": [\n", sizeof(KinCall));
    F(KinCall, precision); F(KinCall, layout); F(KinCall, n); F(KinCall, batch_stride); F(KinCall, q);
    F(KinCall, n_fk_links); F(KinCall, fk_links); F(KinCall, T_out); F(KinCall, n_jac_links); F(KinCall, jac_links);
    F(KinCall, with_rot); F(KinCall, rpy_jac); F(KinCall, keep_irrelevant); F(KinCall, J_out); F(KinCall, truncation_dist);
    F(KinCall, grad_mode); F(KinCall, scratch_mode); F(KinCall, vals_out); F(KinCall, grads_out); F(KinCall, argmin_out);
    F(KinCall, vals_offset); F(KinCall, stream);
    first = 1;
    printf("\n  ]},\n  \"KinIkCall\": {\"size\": %zu, \"fields\": [\n", sizeof(KinIkCall));
    F(KinIkCall, n); F(KinIkCall, link_id); F(KinIkCall, with_rot); F(KinIkCall, iters); F(KinIkCall, ftol); F(KinIkCall, lambda0);
    F(KinIkCall, targets); F(KinIkCall, q0); F(KinIkCall, lower); F(KinIkCall, upper); F(KinIkCall, q_out); F(KinIkCall, f_out);
    F(KinIkCall, iters_out); F(KinIkCall, stream); F(KinIkCall, collision); F(KinIkCall, reserved_); F(KinIkCall, margin);
    F(KinIkCall, coll_weight); F(KinIkCall, ctol); F(KinIkCall, dmin_out);
    printf("\n  ]}\n}\n");
    return 0;
}

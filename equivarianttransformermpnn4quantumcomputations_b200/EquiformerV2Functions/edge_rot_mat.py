"""init_edge_rot_mat (reference edge_rot_mat.py:13-80): per-edge frame whose middle row is the
edge direction.  The helper vector is a torch RNG draw (`torch.rand_like`, reference line 28) and
stays in host-side torch code on purpose: parity needs the *same* draw as the reference
(SURVEY §0.7).  The result is detached -- no gradient flows through the frames."""
import torch


def init_edge_rot_mat(edge_distance_vec):
    vec = edge_distance_vec
    length = vec.pow(2).sum(dim=1).sqrt()
    if torch.min(length) < 0.0001:
        print("Error edge_vec_0_distance: {}".format(torch.min(length)))
    ex = vec / length.view(-1, 1)

    helper = torch.rand_like(vec) - 0.5
    helper = helper / helper.pow(2).sum(dim=1).sqrt().view(-1, 1)
    # two 90-degree alternates in case the draw is (anti)parallel to the edge
    alt_b = torch.stack([-helper[:, 1], helper[:, 0], helper[:, 2]], dim=1)
    alt_c = torch.stack([helper[:, 0], -helper[:, 2], helper[:, 1]], dim=1)

    def absdot(a):
        return (a * ex).sum(dim=1).abs().view(-1, 1)

    dot_b, dot_c = absdot(alt_b), absdot(alt_c)
    helper = torch.where(absdot(helper) > dot_b, alt_b, helper)
    helper = torch.where(absdot(helper) > dot_c, alt_c, helper)
    assert torch.max(absdot(helper)) < 0.99

    ez = torch.cross(ex, helper, dim=1)
    ez = ez / ez.pow(2).sum(dim=1, keepdim=True).sqrt()
    ez = ez / ez.pow(2).sum(dim=1).sqrt().view(-1, 1)
    ey = torch.cross(ex, ez, dim=1)
    ey = ey / ey.pow(2).sum(dim=1, keepdim=True).sqrt()
    # rows of the result: z, x_edge, -y
    return torch.stack([ez, ex, -ey], dim=1).detach()

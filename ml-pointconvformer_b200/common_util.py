"""Hot-path helpers of the reference's util/common_util.py: compute_knn_inverse (250-327),
replace_batchnorm (237-247), to_device (the reference's tensor mover)."""
import torch

from . import pcf_cuda
from . import streams as S


def to_device(inputs, non_blocking=True):
    if isinstance(inputs, (list, tuple)):
        return [to_device(x, non_blocking) for x in inputs]
    return inputs.cuda(non_blocking=non_blocking) if isinstance(inputs, torch.Tensor) else inputs


def compute_knn_inverse(pointclouds, edges_self, edges_forward, edges_propagate):
    """util/common_util.py:250-327 -> (inv_self, inv_forward, inv_propagate), each
    [list_inv_neighbors, list_inv_k, list_inv_idx] (int32 [1,N*K], uint8 [1,N*K], int32 [1,total+1]).
    total_points = pointclouds[j].shape[1] for all three kinds, exactly as the reference does (for
    propagate edges that is the dense level: harmless padding, lines 303-306)."""
    slot = [0]

    def run(edge_list):
        # the 13 edge sets are independent: each inverse map (4 small launches) is built on a side stream
        branches = []
        for j, e in enumerate(edge_list):
            branches.append(S.fork(lambda e=e, j=j: pcf_cuda.compute_knn_inverse(to_device(e).contiguous(), pointclouds[j].shape[1]),
                                   slot[0]))
            slot[0] += 1
        return branches

    def collect(branches):
        out = ([], [], [])
        for b in branches:
            for lst, r in zip(out, S.join(b)):
                lst.append(r)
        return [out[0], out[1], out[2]]
    bs, bf, bp = run(edges_self), run(edges_forward), run(edges_propagate)
    return collect(bs), collect(bf), collect(bp)


def replace_batchnorm(net):
    """Fold every Linear_BN into a Linear for inference (util/common_util.py:237-247)."""
    for name, child in net.named_children():
        if hasattr(child, 'fuse'):
            setattr(net, name, child.fuse())
        elif isinstance(child, torch.nn.BatchNorm2d):
            setattr(net, name, torch.nn.Identity())
        else:
            replace_batchnorm(child)

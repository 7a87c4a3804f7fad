"""TEST INFRASTRUCTURE -- torch restatement of the *kernel's* algorithm (Taylor jets forward +
one reverse sweep), used by the `not gpu` tests to show that the closed-form jet/adjoint
formulas implemented in idff_b200/csrc/idff_field.cuh equal the reference's nested
autograd.grad formulation (oracle.Field2D / FieldMD).  Not used by the product."""
import torch

LA = 1.0507009873554805 * 1.6732632423543772   # lambda*alpha
LAM = 1.0507009873554805


def selu_parts(z):
    """s(z), s'(z), E(z)=s''=s''' (0 on the positive branch; ATen picks the negative branch for z<=0)."""
    neg = z <= 0
    e = LA * torch.exp(z)
    s = torch.where(neg, e - LA, LAM * z)
    s1 = torch.where(neg, e, torch.full_like(z, LAM))
    E = torch.where(neg, e, torch.zeros_like(z))
    return s, s1, E


def momentum_jets(layers, x, t, coef, D):
    """Returns (pred, grad_x[c1*phi + c2*D_d phi + c3*D_d^2 phi]) with phi = sum_k pred_k and
    d = ones over the D x-coordinates.  layers = [(W,b)]*4, first layer input = [x, t, (const)]."""
    order = len(coef)
    c = list(coef) + [0.0] * (3 - order)
    W0, b0 = layers[0]
    N = x.shape[0]
    inp = torch.cat([x, t * torch.ones(N, 1, dtype=x.dtype)], dim=-1)
    z = inp @ W0[:, :inp.shape[1]].T + b0
    zd = W0[:, :D].sum(1).expand_as(z)          # tangent of z1: W0 d
    zdd = torch.zeros_like(z)
    saved = []
    for li in range(3):
        s, s1, E = selu_parts(z)
        a, ad, add = s, s1 * zd, E * zd * zd + s1 * zdd
        saved.append((s1, E, zd, zdd))
        W, b = layers[li + 1]
        if li < 2:
            z, zd, zdd = a @ W.T + b, ad @ W.T, add @ W.T
        else:
            pred = a @ W.T + b
    u3 = layers[3][0].sum(0)                     # adjoint of a3 from phi = 1^T W3 a3
    A0, A1, A2 = c[0] * u3.expand(N, -1), c[1] * u3.expand(N, -1), c[2] * u3.expand(N, -1)
    for li in (2, 1, 0):
        s1, E, zd, zdd = saved[li]
        p = A0 * s1 + A1 * E * zd + A2 * E * (zd * zd + zdd)
        q = A1 * s1 + A2 * 2 * E * zd
        r = A2 * s1
        W = layers[li][0]
        if li > 0:
            A0, A1, A2 = p @ W, q @ W, r @ W
        else:
            g = p @ W[:, :D]
    return pred, g

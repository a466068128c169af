"""TEST INFRASTRUCTURE: a backend for languagegroundedsemseg_b200.step.StepProgram that evaluates every node on the CPU with
the oracle's arithmetic (oracle/me_cpu.py) and obtains each node's backward from torch autograd on that node alone.  Running
the whole program with it and comparing with plain autograd through the same oracle network checks the program's data
flow (which tensor feeds which node, where gradients are accumulated, cat splits) numerically without a GPU."""
import torch

from oracle import me_cpu


class OracleBackend:
    def begin(self, st):
        self.mgr = st.coordinate_manager
        return st.F.detach(), st.coordinate_map_key

    def maps(self, conv, key):
        if conv.use_mm:
            return key, None
        kg = conv.kernel_generator
        if conv.TRANSPOSE:
            out_key = self.mgr.key_with_stride([t // s for t, s in zip(key.tensor_stride, kg.kernel_stride)])
        else:
            out_key = self.mgr.stride(key, kg.kernel_stride) if any(s > 1 for s in kg.kernel_stride) else key
        return out_key, (self.mgr.kernel_map(key, out_key, kg.kernel_size, kg.kernel_dilation, conv.TRANSPOSE),
                         self.mgr.size(out_key))

    def conv_fwd(self, x, conv, km, need_dgrad):
        xr = x.detach().requires_grad_(True)
        with torch.enable_grad():
            if km is None:
                y = xr @ conv.kernel
                if conv.bias is not None:
                    y = y + conv.bias
            else:
                y = me_cpu.sparse_conv(xr, conv.kernel, km[0], km[1], conv.bias)
        return y.detach(), (xr, y, conv)

    def conv_bwd(self, saved, gout, need_gin, need_gb=False):
        xr, y, conv = saved
        wrt = [xr, conv.kernel] + ([conv.bias] if need_gb else [])
        g = torch.autograd.grad(y, wrt, gout)
        return (g[0] if need_gin else None), g[1], (g[2] if need_gb else None)

    def bn_fwd(self, y, res, bn, relu):
        yr = y.detach().requires_grad_(True)
        rr = res.detach().requires_grad_(True) if res is not None else None
        with torch.enable_grad():
            z = bn(yr)
            if rr is not None:
                z = z + rr
            if relu:
                z = torch.relu(z)
        return z.detach(), (yr, rr, z, bn)

    def bn_bwd(self, saved, dz, need_dres):
        yr, rr, z, bn = saved
        wrt = [yr, bn.weight, bn.bias] + ([rr] if rr is not None else [])
        g = torch.autograd.grad(z, wrt, dz)
        return g[0], (g[3] if rr is not None else None), g[1], g[2]

    def ce(self, logits, labels, ignore_index):
        lr = logits.detach().requires_grad_(True)
        with torch.enable_grad():
            loss = torch.nn.functional.cross_entropy(lr, labels.long(), ignore_index=ignore_index)
        return loss.detach(), torch.autograd.grad(loss, lr)[0]

    cat = staticmethod(lambda a, b: torch.cat([a, b], 1))
    add = staticmethod(lambda a, b: a + b)

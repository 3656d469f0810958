"""Reference-shaped ``GPTQ`` / ``Quantizer`` objects backed by the CPU oracle (TEST INFRASTRUCTURE ONLY).

Lets ``gptq_b200.sequential.quantize_sequential`` run the reference algorithm (oracle/gptq_oracle.py,
pinned to the reference's own code) on the host CPU against the very same model activations the CUDA
path sees: the model forward stays on the GPU, every GPTQ computation happens here in torch-CPU fp32."""
import torch

from . import gptq_oracle as O


class OracleQuantizer:
    def __init__(self):
        self.qp = None
        self.maxq = torch.tensor(0)
        self.scale = torch.zeros(1)
        self.zero = torch.zeros(1)

    def configure(self, bits, perchannel=False, sym=True, mse=False, norm=2.4, grid=100, maxshrink=.8, trits=False):
        self.qp = O.QParams(bits=bits, perchannel=perchannel, sym=sym, mse=mse, norm=norm, grid=grid, maxshrink=maxshrink, trits=trits)
        self.maxq = torch.tensor(self.qp.maxq)

    def find_params(self, x, weight=True):
        self.scale, self.zero = O.find_params(x.detach().float().cpu(), self.qp, weight=weight)
        self.qp.scale, self.qp.zero = self.scale, self.zero


def oracle_quantize(x, scale, zero, maxq):
    dev = x.device
    return O.quantize(x.detach().float().cpu(), scale.cpu(), zero.cpu(), int(maxq)).to(dev)


class OracleGPTQ:
    def __init__(self, layer):
        self.layer = layer
        self.dev = layer.weight.device
        self.rows, self.columns = layer.weight.shape
        self.st = O.HessianState(self.columns)
        self.quantizer = None
        self.nsamples = 0

    def add_batch(self, inp, out):
        O.add_batch(self.st, inp.detach().cpu())
        self.nsamples = self.st.nsamples

    @property
    def H(self):
        return self.st.H

    def fasterquant(self, blocksize=128, percdamp=.01, groupsize=-1, actorder=False, static_groups=False, **_):
        W = self.layer.weight.data
        r = O.fasterquant(W.detach().cpu(), self.st.H, self.quantizer.qp, blocksize=blocksize, percdamp=percdamp,
                          groupsize=groupsize, actorder=actorder, static_groups=static_groups, keep_margin=True,
                          stable_perm=True)
        self.oracle_result = r                     # for the rounding-tie classifier of the tests
        self.oracle_diag = torch.diag(self.st.H).clone()
        self.layer.weight.data = r.Q.to(W.dtype).to(W.device)
        self.codes, self.scales, self.zeros, self.g_idx, self.error = r.codes, r.scales, r.zeros, r.g_idx, r.error
        self.quantizer.scale, self.quantizer.zero = r.final_scale, r.final_zero

    def free(self):
        self.st = None

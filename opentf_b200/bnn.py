"""`Bnn`: OpeNTF's Bayesian (Flipout) variant of Fnn (reference: src/mdl/bnn.py) on B200 kernels.

The reference converts every `nn.Linear` of Fnn to bayesian-torch 0.5.0's `LinearFlipout` (bnn.py:19-25): parameters
`mu_weight, rho_weight, mu_bias, rho_bias` per layer, prior N(0,1), posterior init mu ~ N(0, .1), rho ~ N(-3, .1);
one training step draws ONE weight perturbation per layer (fnn.py:126), adds KL/B to the loss (fnn.py:136,149) and
test time averages `nmc` stochastic passes (fnn.py:202-209).  bayesian-torch is not vendored in the reference tree, so
the arithmetic follows its published algorithm -- "parity unpinned" beyond layout and statistics.

Same knobs as the reference block `bnn:` of mdl/__config__.yaml (Fnn's + `nmc`); same files, state_dict keys
`layers.{i}.{mu,rho}_{weight,bias}` in torch [out,in] layout.
"""
import logging

from .fnn import Fnn
from .ntf import Ntf

log = logging.getLogger(__name__)


class Bnn(Fnn):
    def __init__(self, output, device, seed, cfg):
        super().__init__(output, device, seed, cfg)
        self.is_bayesian = True

    @classmethod
    def is_bayesian_cls(cls): return True

    def _host_init(self, input_size, output_size):
        """bnn.py:25: `dnn_to_bnn(super().init(...))` -- the Fnn init is drawn first (and discarded: moped_enable False), then
        every Linear is replaced, in layer order, by a LinearFlipout whose init draws mu_weight, rho_weight, mu_bias, rho_bias."""
        torch = Ntf.torch
        super()._host_init(input_size, output_size)  # consumes the generator exactly as the reference does
        sizes = [input_size] + list(self._c('h')) + [output_size]
        sd = {}
        for i, (fin, fout) in enumerate(zip(sizes[:-1], sizes[1:])):
            sd[f'layers.{i}.mu_weight'] = torch.empty(fout, fin).normal_(0.0, 0.1)
            sd[f'layers.{i}.rho_weight'] = torch.empty(fout, fin).normal_(-3.0, 0.1)
            sd[f'layers.{i}.mu_bias'] = torch.empty(fout).normal_(0.0, 0.1)
            sd[f'layers.{i}.rho_bias'] = torch.empty(fout).normal_(-3.0, 0.1)
        return sd

    def _load_ckpt(self, path):
        sd = super()._load_ckpt(path)
        # bayesian-torch registers its noise / prior tensors as buffers (eps_weight, prior_weight_mu, ...): they ride along in
        # reference checkpoints and carry no state
        return {k: v for k, v in sd.items() if k.split('.')[-1] in ('mu_weight', 'rho_weight', 'mu_bias', 'rho_bias')}

"""`Ntf`: the model contract of OpeNTF (reference: src/mdl/ntf.py:5-134) for the B200 path.

Same constructor, `name()`, `learn / test / evaluate / adila` entry points, run-directory naming and file
formats, so `models.instances=[opentf_b200.fnn.Fnn]` drops into OpeNTF's `main.py` dispatch (main.py:155-181).
"""
import logging
import os
import pickle
import re

import scipy.sparse

from . import metric, util

log = logging.getLogger(__name__)


class Ntf:
    torch = None

    def __init__(self, output, device, seed, cfg):
        import torch
        Ntf.torch = torch
        self.cfg, self.seed, self.device = cfg, seed, device
        self.model, self.is_bayesian = None, False
        self.writer = util.summary_writer()
        util.set_seed(self.seed, torch)  # ntf.py:14
        self.output = output + self.name()
        os.makedirs(self.output, exist_ok=True)

    def name(self):
        return f'/{self.__class__.__name__.lower()}.{util.cfg2str(self.cfg)}'  # ntf.py:27

    def learn(self, teamsvecs, splits, prev_model): pass

    def test(self, teamsvecs, splits, testcfg): pass

    def evaluate(self, teamsvecs, splits, evalcfg):
        """ntf.py:32-92: per fold and prediction file -> trec metrics (+ aucroc, skill coverage) -> CSVs, fold mean/std."""
        import pandas as pd
        torch = Ntf.torch
        assert os.path.isdir(self.output), f'No folder for {self.output} exist!'
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            # one process per GPU: the CSVs are written once, by rank 0; the others wait so that whatever follows sees complete files
            if torch.distributed.get_rank() != 0:
                torch.distributed.barrier()
                return
            try: return self._evaluate(teamsvecs, splits, evalcfg)
            finally: torch.distributed.barrier()
        return self._evaluate(teamsvecs, splits, evalcfg)

    def _evaluate(self, teamsvecs, splits, evalcfg):
        import pandas as pd
        torch = Ntf.torch
        g = util.cfg_get
        y_test = teamsvecs['member'][splits['test']]
        mcfg = g(evalcfg, 'metrics')
        trec, other = list(g(mcfg, 'trec', []) or []), list(g(mcfg, 'other', []) or [])
        for pred_set in (['test', 'train', 'valid'] if g(evalcfg, 'on_train') else ['test']):
            fold_mean, mean_std, fold_inst = pd.DataFrame(), pd.DataFrame(), pd.DataFrame()
            for foldidx in splits['folds'].keys():
                Y = y_test if pred_set == 'test' else teamsvecs['member'][splits['folds'][foldidx][pred_set]]
                predfiles = [f'{self.output}/f{foldidx}.{pred_set}.pred']
                if g(evalcfg, 'per_epoch'):
                    predfiles += [f'{self.output}/{_}' for _ in os.listdir(self.output) if re.match(rf'f{foldidx}\.{pred_set}\.e\d+\.pred$', _)]
                for i, predfile in enumerate(sorted(sorted(predfiles), key=len)):
                    Y_ = torch.load(predfile, map_location='cpu', weights_only=False)['y_pred']
                    Y_ = util.torch_sparse_2_scipy_sparse(Y_, 'csr') if Y_.is_sparse else Y_.cpu().numpy()
                    assert Y.shape == Y_.shape, f'Shape mismatch between truth Y {Y.shape} vs preds Y_ {Y_.shape}!'
                    df, df_mean = pd.DataFrame(), pd.DataFrame()
                    if trec:
                        # the per-team ranking loop runs on the GPU (ntf_eval_ranked); a non-CUDA device raises there: no host fallback
                        df, df_mean = metric.calculate_metrics_device(Y, Y_, g(evalcfg, 'topK'), g(evalcfg, 'per_instance'), trec,
                                                                      device=util.first_device(self.device))
                        if df is None: df = pd.DataFrame()
                    if (m := [m for m in other if 'aucroc' in m]):
                        aucroc, fpr_tpr = metric.calculate_auc_roc(Y, Y_, curve=(m[0] == 'aucroc+'))
                        if df_mean.empty: df_mean = pd.DataFrame(columns=['mean'])
                        df_mean.loc['aucroc'] = aucroc
                        if fpr_tpr:
                            with open(f'{predfile}.eval.roc.pkl', 'wb') as f: pickle.dump(fpr_tpr, f)
                    if (m := [m for m in other if 'skill_coverage' in m]):
                        X = teamsvecs['skill'] if scipy.sparse.issparse(teamsvecs['skill']) else teamsvecs['original_skill']
                        X = X[splits['test']] if pred_set == 'test' else X[splits['folds'][foldidx][pred_set]]
                        df_skc, df_mean_skc = metric.calculate_skill_coverage(X, Y_, teamsvecs['skillcoverage'], g(evalcfg, 'per_instance'),
                                                                              topks=m[0].replace('skill_coverage_', ''), device=util.first_device(self.device))
                        df = df_skc if df.empty else pd.concat([df.reset_index(drop=True), df_skc.reset_index(drop=True)], axis=1)
                        df_mean = df_mean_skc if df_mean.empty else pd.concat([df_mean, df_mean_skc], axis=0)
                    if g(evalcfg, 'per_instance'): df.to_csv(f'{predfile}.eval.instance.csv', float_format='%.5f', index=False)
                    df_mean.to_csv(f'{predfile}.eval.mean.csv')
                    if i == 0:
                        fold_mean = pd.concat([fold_mean, df_mean], axis=1)
                        if g(evalcfg, 'per_instance'): fold_inst = fold_inst.add(df, fill_value=0)
            mean_std['mean'] = fold_mean.mean(axis=1)
            mean_std['std'] = fold_mean.std(axis=1)
            mean_std.to_csv(f'{self.output}/{pred_set}.pred.eval.mean.csv')
            if g(evalcfg, 'per_instance'):
                fold_inst.truediv(len(splits['folds'].keys())).to_csv(f'{self.output}/{pred_set}.pred.eval.instance_mean.csv', index=False)

    def adila(self, teamsvecs, splits, faircfg):
        """ntf.py:108-134 hands the *.pred files to the Adila submodule (out of scope here); when this class
        is hosted inside OpeNTF the reference implementation is reused as is."""
        try:
            from mdl.ntf import Ntf as _RefNtf  # only importable inside an OpeNTF checkout
        except ImportError as e:
            raise NotImplementedError('adila() needs the OpeNTF checkout (Adila submodule) on sys.path') from e
        return _RefNtf.adila(self, teamsvecs, splits, faircfg)

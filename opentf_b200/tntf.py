"""`tNtf`: OpeNTF's temporal wrapper (reference: src/mdl/tntf.py:7-44) -- streaming fine-tuning over time intervals.

The wrapped model (our `Fnn` / `Bnn`) is trained year by year: each interval gets its own k-fold split and run directory
`{output}/{year}`, and starts from the previous interval's checkpoints (`prev_model`, fnn.py:101).  On this path that is just
repeated `learn(prev_model=...)` on the same device-resident teamsvecs (staged once; every year is a row range of it), so the
wrapper adds no kernels.  The reference's own `mdl.tntf.tNtf` can wrap `opentf_b200.fnn.Fnn` unchanged (it only touches
`.output / .learn / .test / .evaluate / .adila`); this class is the same logic for standalone use.
"""
import logging
import os
import pickle

import numpy as np

from . import util
from .ntf import Ntf

log = logging.getLogger(__name__)


class tNtf(Ntf):
    def __init__(self, output, device, seed, cfg, model, year_idx):
        super().__init__(output, device, seed, cfg)
        self.model = model
        self.year_idx = year_idx  # [(first team row of the year, year), ...] (team.py: indexes['i2y'])
        self.output = self.model.output

    def name(self): return ''  # tntf.py:15: the run directory is the wrapped model's

    def learn(self, teamsvecs, splits, prev_model):
        from sklearn.model_selection import KFold
        done = [int(item) for item in os.listdir(self.model.output) if item.isdigit()]  # tntf.py:19
        step_ahead, tfolds = int(util.cfg_get(self.cfg, 'step_ahead')), int(util.cfg_get(self.cfg, 'tfolds'))
        for i, v in enumerate(self.year_idx[:-step_ahead]):  # the last years are for test
            if len(done) > 1:  # tntf.py:23-27: resume -- a year counts as trained once a later year's directory exists
                log.info(f'The model has already been trained on year {min(done)}')
                done.remove(min(done))
                continue
            train = np.arange(self.year_idx[i][0], self.year_idx[i + 1][0])
            skf = KFold(n_splits=tfolds, random_state=self.seed, shuffle=True)
            for k, (train_idx, valid_idx) in enumerate(skf.split(train)):
                splits['folds'][k]['train'] = train[train_idx]
                splits['folds'][k]['valid'] = train[valid_idx]
            self.model.output = f'{self.output}/{self.year_idx[i][1]}'
            os.makedirs(self.model.output, exist_ok=True)
            with open(f'{self.model.output}/splits.pkl', 'wb') as f: pickle.dump(splits, f)
            self.model.learn(teamsvecs, splits, prev_model)  # fine-tune over the time intervals, not recursive (tntf.py:37)
            prev_model = {foldidx: f'{self.model.output}/f{foldidx}.pt' for foldidx in splits['folds'].keys()}

    def test(self, teamsvecs, splits, testcfg): self.model.test(teamsvecs, splits, testcfg)

    def evaluate(self, teamsvecs, splits, evalcfg): self.model.evaluate(teamsvecs, splits, evalcfg)

    def adila(self, teamsvecs, splits, faircfg): self.model.adila(teamsvecs, splits, faircfg)

"""Early stopping on the epoch validation loss with the reference's rule (src/mdl/earlystopping.py:26-39): the
first epoch sets the bar; later an epoch is an improvement only if -v_loss >= best + delta, otherwise the
patience counter advances.  The reference constructs it with save_model=False (fnn.py:110), so nothing is saved."""


class EarlyStopping:
    def __init__(self, patience=5, delta=0.0, verbose=False, trace_func=print):
        self.patience, self.delta, self.verbose, self.trace_func = patience, delta, verbose, trace_func
        self.counter, self.best_score, self.early_stop, self.val_loss_min = 0, None, False, float('inf')

    def __call__(self, val_loss, model=None):
        score = -val_loss
        if self.best_score is None or score >= self.best_score + self.delta:
            if self.verbose: self.trace_func(f'Validation loss decreased ({self.val_loss_min:.6f} --> {val_loss:.6f})')
            self.best_score, self.val_loss_min, self.counter = score, val_loss, 0
        else:
            self.counter += 1
            self.trace_func(f'EarlyStopping counter: {self.counter} out of {self.patience}')
            if self.counter >= self.patience: self.early_stop = True
        return self

"""Training log, same interface as the reference's poduqnn/logger.py (Logger :8-85) without TensorFlow.

Provenance: this file is an interface shim OUTSIDE the hot path (out of scope).  Its method
names, attributes and the text of the printed lines are the reference's on purpose -- `train_model` / `fit` callers and
anything that parses the training log keep working -- so it reads like the original; it carries no algorithmic content
and no coverage is claimed for it."""
import time
from datetime import datetime


class Logger:
    def __init__(self, epochs, frequency, silent=False):
        self.start_time = time.time()
        self.prev_time = self.start_time
        self.tf_epochs = epochs
        self.frequency = frequency
        self.silent = silent
        self.epochs = []
        self.logs = []
        self.logs_keys = None
        self.get_val_err = None

    def get_elapsed(self):
        return datetime.fromtimestamp(time.time() - self.start_time).strftime("%M:%S")

    def set_val_err_fn(self, fn):
        self.get_val_err = fn

    def log_train_start(self):
        if not self.silent:
            print("\nTraining started")
            print("================")

    def log_train_epoch(self, epoch, loss, custom="", is_iter=False):
        """logger.py:45-64: a line every `frequency` epochs: loss + the validation errors."""
        if epoch % self.frequency != 0 or self.silent:
            return
        logs = {"L": loss, **(self.get_val_err() if self.get_val_err else {})}
        if self.logs_keys is None:
            self.logs_keys = list(logs.keys())
        values = [logs[k] for k in self.logs_keys]
        self.epochs.append(epoch)
        self.logs.append(values)
        msg = "".join(f" {k}: {x:.4f}" if i >= 3 else f" {k}: {x:.4e}"
                      for i, (k, x) in enumerate(zip(self.logs_keys, values)))
        print(f"{'nt_epoch' if is_iter else '#'}: {epoch:6d} " + msg + " " + custom)

    def log_train_end(self, epoch, loss, custom=""):
        if self.silent:
            return
        self.log_train_epoch(epoch, loss, custom)
        print("==================")
        print(f"Training finished (epoch {epoch}): duration = {self.get_elapsed()}  " + custom)

    def get_logs(self):
        """(header, array of [epoch, values...]) -- the reference's version returns None (its body is commented
        out, logger.py:75-85); the logged values are kept here."""
        if self.silent or self.logs_keys is None:
            return None
        import numpy as np
        return ("epoch\t" + "\t".join(self.logs_keys),
                np.hstack((np.array(self.epochs)[:, None], np.array(self.logs, dtype=float))))

"""The fit loop's host side on the B200 path (the reference's l3embedding/train.py).

Public names, arguments, output files and the resume rule are the reference's (train.py:29-421):
    LossHistory :29, TimeHistory :108, data_generator :142, single_epoch_data_generator :198, get_restart_info :208,
    train :218 -- plus the keras callbacks train() instantiates (ModelCheckpoint, CSVLogger; keras 2.0.9 semantics).
What is different is how the work is organised:

* batches come from `pipeline.BatchPlan` + `pipeline.PinnedBatchReader` (a reader thread filling page-locked uint8 /
  int16 buffers ahead of the step; resume = a seek in the plan), not from a concatenating loop in the training thread;
  they are RAW by default (`scale_on_host=False`): `2*(u8/255)-1` (train.py:186) and `pcm2float` (train.py:189) run on
  the device and the H2D copy is 4x smaller.  `scale_on_host=True` yields the reference's float arrays;
* under data parallelism (one process per GPU, torchrun) only rank 0 creates the model directory and writes files;
  BN moving statistics are averaged over the ranks before validation / checkpointing and validation is sharded, so all
  ranks see the same epoch logs and take the same save-best decisions;
* checkpoints stay keras-layout weight files (what the reference's tooling loads), and next to every one of them the
  Adam state (step count + both moment arenas) is saved as `<name>.optimizer.npz`; resume restores it when present
  (the reference restarts Adam from zero on resume: `save_weights_only=True`, train.py:329-352).
Google-Sheets logging (train.py:55-105) is out of scope: the gsheet arguments are accepted and ignored.
"""
from __future__ import annotations

import csv
import datetime
import getpass
import json
import logging
import os
import pickle
import time

import numpy as np

from . import pipeline
from .audio import pcm2float
from .model import MODELS, Adam, load_model

LOGGER = logging.getLogger('l3embedding')
LOGGER.setLevel(logging.DEBUG)


# ---- callbacks -----------------------------------------------------------------------------------------------------

class Callback:
    def __init__(self):
        self.model = None

    def set_model(self, model):
        self.model = model


class LossHistory(Callback):
    """Keras callback to record loss history (train.py:29-53)."""

    def __init__(self, outfile):
        super().__init__()
        self.outfile = outfile

    def on_train_begin(self, logs=None):
        self.loss = []
        self.val_loss = []

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        self.loss.append(logs.get('loss'))
        self.val_loss.append(logs.get('val_loss'))
        with open(self.outfile, 'wb') as fp:
            pickle.dump({'loss': self.loss, 'val_loss': self.val_loss}, fp)


class TimeHistory(Callback):
    """Keras callback to log epoch and batch running time (train.py:108-131)."""

    def on_train_begin(self, logs=None):
        self.epoch_times = []
        self.batch_times = []

    def on_epoch_begin(self, batch, logs=None):
        self.epoch_time_start = time.time()

    def on_epoch_end(self, batch, logs=None):
        t = time.time() - self.epoch_time_start
        LOGGER.info('Epoch took {} seconds'.format(t))
        self.epoch_times.append(t)

    def on_batch_begin(self, batch, logs=None):
        self.batch_time_start = time.time()

    def on_batch_end(self, batch, logs=None):
        t = time.time() - self.batch_time_start
        LOGGER.debug('Batch took {} seconds'.format(t))
        self.batch_times.append(t)


OPTIMIZER_SUFFIX = '.optimizer.npz'


def optimizer_state_path(weights_path):
    return os.path.splitext(weights_path)[0] + OPTIMIZER_SUFFIX


class ModelCheckpoint(Callback):
    """keras.callbacks.ModelCheckpoint as train.py:327-355 uses it (save_weights_only, save_best_only, period,
    `{epoch:02d}` in the path, monitor 'val_acc' -> max / 'val_loss' -> min).  save_optimizer additionally writes the
    Adam state next to the weight file (see the module docstring)."""

    def __init__(self, filepath, monitor='val_loss', verbose=0, save_best_only=False, save_weights_only=True, period=1,
                 save_optimizer=True):
        super().__init__()
        self.filepath, self.monitor, self.verbose = filepath, monitor, verbose
        self.save_best_only, self.period = save_best_only, period
        self.save_optimizer = save_optimizer
        self.epochs_since_last_save = 0
        self.maximize = 'acc' in monitor or monitor.startswith('fmeasure')
        self.best = -np.inf if self.maximize else np.inf

    def _save(self, path):
        self.model.save_weights(path)
        state = self.model.get_optimizer_state() if self.save_optimizer and hasattr(self.model, 'get_optimizer_state') else None
        if state is not None:
            with open(optimizer_state_path(path), 'wb') as f:
                np.savez(f, t=np.int64(state['t']), m=state['m'], v=state['v'])

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        self.epochs_since_last_save += 1
        if self.epochs_since_last_save < self.period:
            return
        self.epochs_since_last_save = 0
        path = self.filepath.format(epoch=epoch + 1, **logs)
        if self.save_best_only:
            cur = logs.get(self.monitor)
            if cur is None:
                return
            if (cur > self.best) if self.maximize else (cur < self.best):
                self.best = cur
                self._save(path)
        else:
            self._save(path)


class CSVLogger(Callback):
    """keras.callbacks.CSVLogger(path, append=True, separator=','): columns `epoch` + sorted log keys -- the file
    get_restart_info and 04_plot_training_history.py read."""

    def __init__(self, filename, separator=',', append=False):
        super().__init__()
        self.filename, self.sep, self.append = filename, separator, append
        self.keys = None

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        if self.keys is None:
            self.keys = sorted(logs.keys())
            exists = os.path.exists(self.filename) and os.path.getsize(self.filename) > 0
            self._header = not (self.append and exists)
            if not self.append and exists:
                os.remove(self.filename)
        with open(self.filename, 'a', newline='') as f:
            w = csv.DictWriter(f, fieldnames=['epoch'] + self.keys, delimiter=self.sep)
            if self._header:
                w.writeheader()
                self._header = False
            row = {'epoch': epoch}
            row.update({k: logs[k] for k in self.keys})
            w.writerow(row)


# ---- batches -------------------------------------------------------------------------------------------------------

def _to_reference_floats(batch):
    """train.py:186,189: what the reference's generator yields."""
    out = dict(batch)
    if 'video' in out:
        out['video'] = 2 * (np.asarray(out['video']).astype('float64') / 255.0).astype('float32') - 1
    if 'audio' in out:
        out['audio'] = pcm2float(np.asarray(out['audio']), dtype='float32')
    return out


def data_generator(data_dir, batch_size=512, random_state=20180123, start_batch_idx=None, keys=None,
                   scale_on_host=False, prefetch=True):
    """The batch stream of train.py:142-195 (same sequence, same resume rule) as dicts with `keys`
    (default audio / video / label).  prefetch=True reads ahead in a thread into a ring of pinned buffers: an item is
    valid until two further items have been taken (copy it if you keep it longer); prefetch=False returns fresh
    arrays.  scale_on_host=True converts to the reference's float arrays (always fresh arrays)."""
    keys = tuple(keys) if keys else pipeline.KEYS
    plan = pipeline.BatchPlan(data_dir, batch_size, random_state)
    start = int(start_batch_idx or 0)
    if prefetch and set(keys) <= set(pipeline.KEYS):
        src = pipeline.PinnedBatchReader(plan, start, keys=keys)
    else:
        src = pipeline.read_batches(plan, start, keys=keys)
    try:
        for batch in src:
            yield _to_reference_floats(batch) if scale_on_host else batch
    finally:
        if hasattr(src, 'close'):
            src.close()


def single_epoch_data_generator(data_dir, epoch_size, **kwargs):
    """train.py:198-205: the first `epoch_size` batches of the stream, over and over (the validation set)."""
    while True:
        gen = data_generator(data_dir, **kwargs)
        try:
            for _ in range(epoch_size):
                yield next(gen)
        finally:
            gen.close()


def keras_tuples(gen, inputs, outputs):
    """pescador.maps.keras_tuples for the one case train.py:382-395 needs."""
    for item in gen:
        yield [item[k] for k in inputs], item[outputs]


def get_restart_info(history_path):
    """(last epoch index, its val_acc, its val_loss) from the CSV history (train.py:208-215)."""
    with open(history_path, 'r') as f:
        rows = list(csv.DictReader(f))
    last = rows[-1]
    return int(last['epoch']), float(last['val_acc']), float(last['val_loss'])


# ---- train() -------------------------------------------------------------------------------------------------------

def _replicas(gpus):
    """(rank, world size, broadcast(obj) -> obj) of the data-parallel job this process belongs to."""
    if gpus is None or gpus <= 1:
        return 0, 1, (lambda obj: obj)
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("gpus=%d needs one process per GPU: launch with torchrun and call "
                           "torch.distributed.init_process_group first" % gpus)

    def bcast(obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]
    return dist.get_rank(), dist.get_world_size(), bcast


def _model_id(train_data_dir, model_type):
    subset = os.path.basename(os.path.normpath(train_data_dir))
    if '_' in subset:
        subset = subset[:subset.rindex('_')]          # train.py:233-234
    return os.path.join(subset, model_type)


def _file_callbacks(model_dir, checkpoint_interval, restart):
    """The file-writing callbacks of train.py:327-365 (rank 0 only)."""
    latest = ModelCheckpoint(os.path.join(model_dir, 'model_latest.h5'), save_weights_only=True, verbose=1)
    best_acc = ModelCheckpoint(os.path.join(model_dir, 'model_best_valid_accuracy.h5'), save_weights_only=True,
                               save_best_only=True, verbose=1, monitor='val_acc')
    best_loss = ModelCheckpoint(os.path.join(model_dir, 'model_best_valid_loss.h5'), save_weights_only=True,
                                save_best_only=True, verbose=1, monitor='val_loss')
    periodic = ModelCheckpoint(os.path.join(model_dir, 'model_checkpoint.{epoch:02d}.h5'), save_weights_only=True,
                               period=checkpoint_interval)
    if restart is not None:
        last_epoch_idx, last_val_acc, last_val_loss = restart
        best_acc.best = last_val_acc
        best_loss.best = last_val_loss
        periodic.epochs_since_last_save = (last_epoch_idx + 1) % checkpoint_interval
    return [latest, best_acc, best_loss, periodic, TimeHistory(),
            LossHistory(os.path.join(model_dir, 'history_checkpoint.pkl')),
            CSVLogger(os.path.join(model_dir, 'history_csvlog.csv'), append=True, separator=',')]


def train(train_data_dir, validation_data_dir, output_dir,
          num_epochs=150, train_epoch_size=512, validation_epoch_size=1024,
          train_batch_size=64, validation_batch_size=64,
          model_type='cnn_L3_orig', random_state=20180123,
          learning_rate=1e-4, verbose=False, checkpoint_interval=10,
          log_path=None, disable_logging=False, gpus=1, continue_model_dir=None,
          gsheet_id=None, google_dev_app_name=None, dtype=None, scale_on_host=False):
    """train.py:218-421 with the same arguments (gsheet_* ignored), file outputs and resume behaviour.
    Returns (model_dir, history)."""
    rank, world, bcast = _replicas(gpus)
    chief = rank == 0
    model_id = _model_id(train_data_dir, model_type)
    param_dict = {
        'username': getpass.getuser(), 'train_data_dir': train_data_dir, 'validation_data_dir': validation_data_dir,
        'model_id': model_id, 'output_dir': output_dir, 'num_epochs': num_epochs, 'train_epoch_size': train_epoch_size,
        'validation_epoch_size': validation_epoch_size, 'train_batch_size': train_batch_size,
        'validation_batch_size': validation_batch_size, 'model_type': model_type, 'random_state': random_state,
        'learning_rate': learning_rate, 'verbose': verbose, 'checkpoint_interval': checkpoint_interval,
        'log_path': log_path, 'disable_logging': disable_logging, 'gpus': gpus, 'continue_model_dir': continue_model_dir,
        'gsheet_id': gsheet_id, 'google_dev_app_name': google_dev_app_name,
    }
    if chief:
        LOGGER.info('Training with the following arguments: {}'.format(param_dict))

    restart = None
    if continue_model_dir:
        latest = os.path.join(continue_model_dir, 'model_latest.h5')
        m, inputs, outputs = load_model(latest, model_type, return_io=True, src_num_gpus=gpus)
        restart = get_restart_info(os.path.join(continue_model_dir, 'history_csvlog.csv'))
        opt_path = optimizer_state_path(latest)
        if os.path.exists(opt_path):      # written by this package; the reference's checkpoints carry no optimizer state
            with np.load(opt_path) as z:
                m.set_optimizer_state({'t': int(z['t']), 'm': z['m'], 'v': z['v']})
        model_dir = continue_model_dir
    else:
        m, inputs, outputs = MODELS[model_type](num_gpus=gpus)
        # one directory for the whole job: rank 0 names it (the ranks' clocks may straddle a second)
        model_dir = bcast(os.path.join(output_dir, 'embedding', model_id,
                                       datetime.datetime.now().strftime("%Y%m%d%H%M%S")) if chief else None)
    if dtype:
        m.configure(dtype=dtype)
    m.compile(Adam(lr=learning_rate), loss='categorical_crossentropy', metrics=['accuracy'])

    callbacks = []
    if chief:
        os.makedirs(model_dir, exist_ok=True)
        param_dict['model_dir'] = model_dir
        with open(os.path.join(model_dir, 'config.json'), 'w') as fd:
            json.dump(param_dict, fd, indent=2)
        with open(os.path.join(model_dir, 'model_spec.pkl'), 'wb') as fd:
            pickle.dump(m.get_config(), fd)
        with open(os.path.join(model_dir, 'model.json'), 'w') as fd:
            json.dump(m.to_json(), fd, indent=2)
        callbacks = _file_callbacks(model_dir, checkpoint_interval, restart)

    first_epoch = restart[0] + 1 if restart is not None else 0
    train_gen = keras_tuples(data_generator(train_data_dir, batch_size=train_batch_size, random_state=random_state,
                                            start_batch_idx=(train_epoch_size * first_epoch) if restart is not None else None,
                                            scale_on_host=scale_on_host),
                             ['video', 'audio'], 'label')
    val_gen = keras_tuples(single_epoch_data_generator(validation_data_dir, validation_epoch_size,
                                                       batch_size=validation_batch_size, random_state=random_state,
                                                       scale_on_host=scale_on_host),
                           ['video', 'audio'], 'label')
    try:
        history = m.fit_generator(train_gen, train_epoch_size, num_epochs, validation_data=val_gen,
                                  validation_steps=validation_epoch_size, callbacks=callbacks,
                                  verbose=1 if (verbose and chief) else 0, initial_epoch=first_epoch)
    finally:
        train_gen.close()
        val_gen.close()
    if chief:
        with open(os.path.join(model_dir, 'history.pkl'), 'wb') as fd:
            pickle.dump(history.history, fd)
    return model_dir, history

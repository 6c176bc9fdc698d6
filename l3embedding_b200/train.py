"""l3embedding/train.py on the B200 path: the fit loop's host side.

Mirrors the reference's names, arguments, file outputs and resume rule (train.py:29-421):
    LossHistory :29, TimeHistory :108, cycle_shuffle :134, data_generator :142, single_epoch_data_generator :198,
    get_restart_info :208, train :218
plus the two keras callbacks train() instantiates (ModelCheckpoint, CSVLogger; keras 2.0.9 semantics).  The heavy
imports the reference pulls in at module scope (git, gsheets, pescador, skimage, h5py, googleapiclient) are gone
or lazy; Google-Sheets logging is out of scope (gsheet arguments are accepted and ignored).

Differences that matter for speed, not results: batches are yielded as RAW uint8 video / int16 audio by default
(`scale_on_host=False`) -- `2*(u8/255)-1` (train.py:186) and `pcm2float` (train.py:189) run on the device and the H2D
copy is 4x smaller; pass scale_on_host=True for the reference's float arrays (bit-identical model inputs).
Batch files: the reference's gzip HDF5 blobs (data/avc/sample.py:565-568) are read with h5py when it exists and with
the built-in reader (l3embedding_b200/minihdf5.py: chunked + deflate/shuffle datasets) otherwise; `.npz` files with the
same three keys (`audio` (n,1,48000) int16, `video` (n,224,224,3) uint8, `label` (n,2)) work too.
"""
from __future__ import annotations

import csv
import datetime
import getpass
import json
import logging
import os
import pickle
import random
import time

import numpy as np

from .audio import pcm2float
from .model import MODELS, Adam, load_model

LOGGER = logging.getLogger('l3embedding')
LOGGER.setLevel(logging.DEBUG)


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


class ModelCheckpoint(Callback):
    """keras.callbacks.ModelCheckpoint as train.py:327-355 uses it (save_weights_only, save_best_only, period,
    `{epoch:02d}` in the path, monitor 'val_acc' -> max / 'val_loss' -> min)."""

    def __init__(self, filepath, monitor='val_loss', verbose=0, save_best_only=False, save_weights_only=True, period=1):
        super().__init__()
        self.filepath, self.monitor, self.verbose = filepath, monitor, verbose
        self.save_best_only, self.period = save_best_only, period
        self.epochs_since_last_save = 0
        self.maximize = 'acc' in monitor or monitor.startswith('fmeasure')
        self.best = -np.inf if self.maximize else np.inf

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
                self.model.save_weights(path)
        else:
            self.model.save_weights(path)


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


def cycle_shuffle(iterable, shuffle=True):
    lst = list(iterable)
    while True:
        yield from lst
        if shuffle:
            random.shuffle(lst)


class _H5Blob:
    """The three datasets of one reference batch file read with the built-in HDF5 reader (minihdf5): each dataset is
    decompressed once, on first use, and sliced from memory afterwards."""

    def __init__(self, path):
        from . import minihdf5
        self._f = minihdf5.File(path)
        self._cache = {}

    def __getitem__(self, key):
        if key not in self._cache:
            self._cache[key] = np.asarray(self._f[key])
        return self._cache[key]

    def close(self):
        self._cache.clear()


def _open_blob(path):
    if path.endswith('.npz'):
        return np.load(path), lambda b: b.close()
    try:
        import h5py  # the reference's reader, when it exists
    except ImportError:
        return _H5Blob(path), lambda b: b.close()
    f = h5py.File(path, 'r')
    return f, lambda b: b.close()


def data_generator(data_dir, batch_size=512, random_state=20180123, start_batch_idx=None, keys=None,
                   scale_on_host=False):
    """train.py:142-195: concatenates slices of the batch files into batches of `batch_size`, skipping (without
    reading) everything before `start_batch_idx` when resuming."""
    random.seed(random_state)
    batch = None
    curr_batch_size = 0
    batch_idx = 0
    if not keys:
        keys = ['audio', 'video', 'label']
    for fname in cycle_shuffle(sorted(os.listdir(data_dir))):
        blob, close = _open_blob(os.path.join(data_dir, fname))
        blob_size = len(blob['label'])
        blob_start_idx = 0
        while blob_start_idx < blob_size:
            blob_end_idx = min(blob_start_idx + batch_size - curr_batch_size, blob_size)
            if start_batch_idx is None or batch_idx >= start_batch_idx:
                if batch is None:
                    batch = {k: blob[k][blob_start_idx:blob_end_idx] for k in keys}
                else:
                    for k in keys:
                        batch[k] = np.concatenate([batch[k], blob[k][blob_start_idx:blob_end_idx]])
            curr_batch_size += blob_end_idx - blob_start_idx
            blob_start_idx = blob_end_idx
            if curr_batch_size == batch_size:
                if start_batch_idx is None or batch_idx >= start_batch_idx:
                    if scale_on_host:
                        # train.py:186,189
                        batch['video'] = 2 * (batch['video'].astype('float64') / 255.0).astype('float32') - 1
                        batch['audio'] = pcm2float(batch['audio'], dtype='float32')
                    batch['label'] = np.asarray(batch['label'], dtype=np.float32)
                    yield batch
                batch_idx += 1
                curr_batch_size = 0
                batch = None
        close(blob)


def single_epoch_data_generator(data_dir, epoch_size, **kwargs):
    while True:
        data_gen = data_generator(data_dir, **kwargs)
        for idx, item in enumerate(data_gen):
            yield item
            if (idx + 1) == epoch_size:
                break


def keras_tuples(gen, inputs, outputs):
    """pescador.maps.keras_tuples for the one case train.py:382-395 needs."""
    for item in gen:
        yield [item[k] for k in inputs], item[outputs]


def get_restart_info(history_path):
    last = None
    with open(history_path, 'r') as f:
        for row in csv.DictReader(f):
            last = row
    return int(last['epoch']), float(last['val_acc']), float(last['val_loss'])


def train(train_data_dir, validation_data_dir, output_dir,
          num_epochs=150, train_epoch_size=512, validation_epoch_size=1024,
          train_batch_size=64, validation_batch_size=64,
          model_type='cnn_L3_orig', random_state=20180123,
          learning_rate=1e-4, verbose=False, checkpoint_interval=10,
          log_path=None, disable_logging=False, gpus=1, continue_model_dir=None,
          gsheet_id=None, google_dev_app_name=None, dtype=None, scale_on_host=False):
    """train.py:218-421 with the same arguments (gsheet_* ignored), file outputs and resume behaviour."""
    data_subset_name = os.path.basename(os.path.normpath(train_data_dir))
    if '_' in data_subset_name:
        data_subset_name = data_subset_name[:data_subset_name.rindex('_')]
    model_id = os.path.join(data_subset_name, model_type)
    param_dict = {
        'username': getpass.getuser(), 'train_data_dir': train_data_dir, 'validation_data_dir': validation_data_dir,
        'model_id': model_id, 'output_dir': output_dir, 'num_epochs': num_epochs, 'train_epoch_size': train_epoch_size,
        'validation_epoch_size': validation_epoch_size, 'train_batch_size': train_batch_size,
        'validation_batch_size': validation_batch_size, 'model_type': model_type, 'random_state': random_state,
        'learning_rate': learning_rate, 'verbose': verbose, 'checkpoint_interval': checkpoint_interval,
        'log_path': log_path, 'disable_logging': disable_logging, 'gpus': gpus, 'continue_model_dir': continue_model_dir,
        'gsheet_id': gsheet_id, 'google_dev_app_name': google_dev_app_name,
    }
    LOGGER.info('Training with the following arguments: {}'.format(param_dict))

    if continue_model_dir:
        m, inputs, outputs = load_model(os.path.join(continue_model_dir, 'model_latest.h5'), model_type, return_io=True,
                                        src_num_gpus=gpus)
    else:
        m, inputs, outputs = MODELS[model_type](num_gpus=gpus)
    if dtype:
        m.configure(dtype=dtype)

    if continue_model_dir:
        model_dir = continue_model_dir
    else:
        model_dir = os.path.join(output_dir, 'embedding', model_id, datetime.datetime.now().strftime("%Y%m%d%H%M%S"))
    os.makedirs(model_dir, exist_ok=True)

    m.compile(Adam(lr=learning_rate), loss='categorical_crossentropy', metrics=['accuracy'])
    param_dict['model_dir'] = model_dir
    with open(os.path.join(model_dir, 'config.json'), 'w') as fd:
        json.dump(param_dict, fd, indent=2)
    with open(os.path.join(model_dir, 'model_spec.pkl'), 'wb') as fd:
        pickle.dump(m.get_config(), fd)
    with open(os.path.join(model_dir, 'model.json'), 'w') as fd:
        json.dump(m.to_json(), fd, indent=2)

    if continue_model_dir is not None:
        last_epoch_idx, last_val_acc, last_val_loss = get_restart_info(os.path.join(continue_model_dir, 'history_csvlog.csv'))

    cb = [ModelCheckpoint(os.path.join(model_dir, 'model_latest.h5'), save_weights_only=True, verbose=1)]
    best_val_acc_cb = ModelCheckpoint(os.path.join(model_dir, 'model_best_valid_accuracy.h5'), save_weights_only=True,
                                      save_best_only=True, verbose=1, monitor='val_acc')
    best_val_loss_cb = ModelCheckpoint(os.path.join(model_dir, 'model_best_valid_loss.h5'), save_weights_only=True,
                                       save_best_only=True, verbose=1, monitor='val_loss')
    checkpoint_cb = ModelCheckpoint(os.path.join(model_dir, 'model_checkpoint.{epoch:02d}.h5'), save_weights_only=True,
                                    period=checkpoint_interval)
    if continue_model_dir is not None:
        best_val_acc_cb.best = last_val_acc
        best_val_loss_cb.best = last_val_loss
        checkpoint_cb.epochs_since_last_save = (last_epoch_idx + 1) % checkpoint_interval
    cb += [best_val_acc_cb, best_val_loss_cb, checkpoint_cb, TimeHistory(),
           LossHistory(os.path.join(model_dir, 'history_checkpoint.pkl')),
           CSVLogger(os.path.join(model_dir, 'history_csvlog.csv'), append=True, separator=',')]

    train_start_batch_idx = train_epoch_size * (last_epoch_idx + 1) if continue_model_dir is not None else None
    train_gen = keras_tuples(data_generator(train_data_dir, batch_size=train_batch_size, random_state=random_state,
                                            start_batch_idx=train_start_batch_idx, scale_on_host=scale_on_host),
                             ['video', 'audio'], 'label')
    val_gen = keras_tuples(single_epoch_data_generator(validation_data_dir, validation_epoch_size,
                                                       batch_size=validation_batch_size, random_state=random_state,
                                                       scale_on_host=scale_on_host),
                           ['video', 'audio'], 'label')

    initial_epoch = last_epoch_idx + 1 if continue_model_dir is not None else 0
    history = m.fit_generator(train_gen, train_epoch_size, num_epochs, validation_data=val_gen,
                              validation_steps=validation_epoch_size, callbacks=cb, verbose=1 if verbose else 0,
                              initial_epoch=initial_epoch)
    with open(os.path.join(model_dir, 'history.pkl'), 'wb') as fd:
        pickle.dump(history.history, fd)
    return model_dir, history

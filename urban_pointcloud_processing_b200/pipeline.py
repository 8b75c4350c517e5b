"""Data-fusion pipeline (drop-in for ``upcp.pipeline.Pipeline``, reference
src/upcp/pipeline.py:18-196).

``process_cloud`` keeps the reference semantics -- processors are applied in order,
each one sees ``mask = (labels == 0)`` (pipeline.py:90) and mutates ``labels`` -- but
uploads the cloud ONCE, chains every drop-in processor on the device and downloads
the labels once at the end.  Processors that are not drop-ins (no
``_label_mask_device``) are still accepted: the labels are synchronised to the host
around them and their own ``get_labels`` is called, exactly as the reference does.
LAS file I/O (``process_file`` / ``process_folder``): ``laspy`` when installed, the native
``utils.las_io`` (``.las`` and, through csrc/laz.cu, ``.laz``) otherwise; with drop-in processors only the raw integer
coordinates are uploaded and scaled on the device."""
import logging
import os
import pathlib
import time

import numpy as np

from . import device as dev
from .abstract_processor import PerThread, drop_per_thread_state
from .analysis import analysis_tools
from .utils import las_utils

logger = logging.getLogger(__name__)


class Pipeline:
    """
    Parameters
    ----------
    processors : iterable of type AbstractProcessor
        The processors to apply, in order.
    exclude_labels : list
        List of labels to exclude from processing.
    ahn_reader : AHNReader object
        Pointer to the AHNReader object used in the processors. Required if caching is
        used.
    caching : bool (default: True)
        Enable caching of AHN interpolation data.
    """

    FILE_TYPES = ('.LAS', '.las', '.LAZ', '.laz')

    last_timings = PerThread(list)      # (processor, seconds) of the calling thread's last tile

    def __init__(self, processors=[], exclude_labels=[], ahn_reader=None, caching=True):
        if ahn_reader is None and caching:
            logger.error('An ahn_reader must be specified when caching is enabled.')
            raise ValueError
        self.processors = processors
        self.exclude_labels = exclude_labels
        self.ahn_reader = ahn_reader
        self.caching = caching
        if self.caching:
            self.ahn_reader.set_caching(self.caching)
        self.last_timings = []

    def __getstate__(self):
        return drop_per_thread_state(self)

    def _create_mask(self, mask, labels):
        """Create mask based on `exclude_labels` (dead in the reference too: the mask is
        overwritten per processor, pipeline.py:90)."""
        if mask is None:
            mask = np.ones((len(labels),), dtype=bool)
        for exclude_label in self.exclude_labels:
            mask = mask & (labels != exclude_label)
        return mask

    def process_tile_device(self, tilecode, tile, d_labels, sync_timings=False):
        """Run all processors on a device-resident tile; ``d_labels`` (int32 cuda) is
        updated in place and returned.  This is the hot path the benchmark times."""
        import torch
        self.last_timings = []
        for obj in self.processors:
            if sync_timings:
                torch.cuda.synchronize()
                start = time.time()
            if not hasattr(obj, '_apply_device'):
                raise TypeError(f'{type(obj).__name__} is not a device drop-in processor')
            # mask = labels == 0 (pipeline.py:90); processors that overwrite it anyway do not get one built
            d_mask = None if getattr(obj, 'ignores_mask', False) else dev.mask_build(d_labels, None, base=False, eq=(0,))
            obj._apply_device(tile, d_labels, d_mask, tilecode)
            if sync_timings:
                torch.cuda.synchronize()
                self.last_timings.append((type(obj).__name__, time.time() - start))
        return d_labels

    def process_cloud(self, tilecode, points, labels, mask=None):
        """
        Process a single point cloud.

        Parameters
        ----------
        tilecode : str
            The CycloMedia tile-code for the given pointcloud.
        points : array of shape (n_points, 3)
            The point cloud <x, y, z>.
        labels : array of shape (n_points,)
            All labels as int values (updated in place).
        mask : array of shape (n_points,) with dtype=bool
            Pre-mask (ignored after the first processor, as in the reference).

        Returns
        -------
        The ``labels`` array with the updated labels.
        """
        tile = dev.get_tile(points)
        d_labels = dev.upload_labels(labels, tile.device)
        host_stale = False
        info = logger.isEnabledFor(logging.INFO)
        for obj in self.processors:
            start = time.time()
            if hasattr(obj, '_apply_device'):
                d_mask = dev.mask_build(d_labels, None, base=False, eq=(0,))
                n_before = dev.mask_count(d_mask) if info else 0
                obj._apply_device(tile, d_labels, d_mask, tilecode)
                host_stale = True
                if info:
                    n_after = dev.mask_count(dev.mask_build(d_labels, None, base=False, eq=(0,)))
                    diff = n_before - n_after
            else:
                if host_stale:
                    dev.download_labels(d_labels, labels)
                    host_stale = False
                host_mask = (labels == 0)
                n_before = np.count_nonzero(host_mask)
                labels = obj.get_labels(points, labels, host_mask, tilecode)
                diff = n_before - np.count_nonzero(labels == 0)
                d_labels = dev.upload_labels(labels, tile.device)
            if info:
                duration = time.time() - start
                logger.info(f'Processor finished in {duration:.2f}s, ' + f'{diff} points labelled.')
        if host_stale:
            dev.download_labels(d_labels, labels)
        return labels

    def _run_tile_workers(self, n_tiles, workers, body):
        """Run ``body(i)`` for i in 0..n_tiles-1 on ``workers`` host threads, each with its own CUDA
        stream (and therefore its own scratch buffers, see device.workspace).  Tiles are independent;
        having 2-3 of them in flight lets the device fill the latency-bound phases of one tile
        (frontier growth, union-find linking, host round trips, uploads) with the kernels of another.
        Tiles are handed out from a shared counter.  The first exception of a worker is re-raised."""
        import threading

        import torch
        device = dev.current_device()
        lock = threading.Lock()
        state = {'next': 0}
        errors = []

        def take():
            with lock:
                i = state['next']
                state['next'] = i + 1
            return i

        def run(k):
            try:
                torch.cuda.set_device(device)               # the current device is per host thread
                stream = dev.worker_stream(device, k)       # persistent: its scratch buffers are re-used
                with torch.cuda.stream(stream):
                    while not errors:
                        i = take()
                        if i >= n_tiles:
                            break
                        body(i)
                    stream.synchronize()
            except BaseException as exc:                    # noqa: BLE001 -- handed to the caller
                errors.append(exc)

        threads = [threading.Thread(target=run, args=(k,), name=f'upcp-tile-worker-{k}') for k in range(workers)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]

    def process_tiles_device(self, tilecodes, tiles, d_labels_list, workers=2):
        """``process_tile_device`` for a batch of device-resident tiles, ``workers`` tiles in flight
        (one CUDA stream and host thread each).  Every ``d_labels_list[i]`` is updated in place."""
        n_tiles = len(tilecodes)
        if not (len(tiles) == n_tiles and len(d_labels_list) == n_tiles):
            logger.error('tilecodes, tiles and d_labels_list must have the same length')
            raise ValueError
        workers = max(1, min(int(workers), n_tiles))
        if workers == 1:
            for tc, tile, d_labels in zip(tilecodes, tiles, d_labels_list):
                self.process_tile_device(tc, tile, d_labels)
            return d_labels_list
        import torch
        main = torch.cuda.current_stream()
        ready = main.record_event()                         # whatever produced the inputs on the caller's stream

        def body(i):
            torch.cuda.current_stream().wait_event(ready)
            self.process_tile_device(tilecodes[i], tiles[i], d_labels_list[i])
        self._run_tile_workers(n_tiles, workers, body)       # workers synchronise their streams before returning
        return d_labels_list

    def process_clouds(self, tilecodes, clouds, labels_list, workers=2):
        """Process a batch of point clouds (what ``process_folder`` does file by file,
        pipeline.py:186-194, for clouds that are already in host memory).

        Tiles are independent, so ``workers`` of them are in flight at once, each on its own CUDA
        stream driven by its own host thread: while one tile is being uploaded (truly asynchronous
        when the host arrays are pinned) or sits in a latency-bound phase, the kernels of another
        run.  With ``workers=1`` a single stream is used and only the upload of tile ``i + 1`` is
        overlapped with the kernels of tile ``i`` (second copy stream).  Per tile the result is
        exactly that of ``process_cloud``: every ``labels_list[i]`` is updated in place.  Returns
        ``labels_list``.
        """
        import torch
        n_tiles = len(tilecodes)
        if not (len(clouds) == n_tiles and len(labels_list) == n_tiles):
            logger.error('tilecodes, clouds and labels_list must have the same length')
            raise ValueError
        if not all(hasattr(obj, '_apply_device') for obj in self.processors):
            for tc, pts, lab in zip(tilecodes, clouds, labels_list):
                self.process_cloud(tc, pts, lab)
            return labels_list
        device = dev.current_device()
        workers = max(1, min(int(workers), n_tiles))
        if workers > 1:
            def body(i):
                tile = dev.DeviceTile(clouds[i], device)
                d_labels = dev.upload_labels(labels_list[i], device)
                self.process_tile_device(tilecodes[i], tile, d_labels)
                dev.download_labels(d_labels, labels_list[i])
            self._run_tile_workers(n_tiles, workers, body)
            return labels_list
        main = torch.cuda.current_stream(device)
        side = dev.worker_stream(device, 'copy')

        def prefetch(i):
            # the side stream may recycle memory the main stream has finished with
            side.wait_stream(main)
            with torch.cuda.stream(side):
                tile = dev.DeviceTile(clouds[i], device)
                d_labels = dev.upload_labels(labels_list[i], device)
                done = side.record_event()
            return tile, d_labels, done

        nxt = prefetch(0) if n_tiles else None
        for i in range(n_tiles):
            tile, d_labels, done = nxt
            nxt = prefetch(i + 1) if i + 1 < n_tiles else None      # queued BEFORE the kernels of tile i
            main.wait_event(done)
            self.process_tile_device(tilecodes[i], tile, d_labels)
            dev.download_labels(d_labels, labels_list[i])
            del tile, d_labels
        return labels_list

    def _load_file(self, in_file):
        """Host side of process_file before the processors: read + parse (pipeline.py:114-121)."""
        tilecode = las_utils.get_tilecode_from_filename(in_file)
        pointcloud = las_utils.read_las(in_file)
        if 'label' not in pointcloud.point_format.extra_dimension_names:
            labels = np.zeros((pointcloud.header.point_count,), dtype='uint16')
        else:
            labels = np.asarray(pointcloud.label)
        return tilecode, pointcloud, labels

    def _label_file(self, tilecode, pointcloud, labels, mask=None):
        """The processors on one loaded file; returns the final labels."""
        if all(hasattr(obj, '_apply_device') for obj in self.processors):
            # device path: raw int32 X, Y, Z up (12 B/pt), coordinates rebuilt in fp64 on the device
            tile = dev.DeviceTile.from_las(pointcloud)
            d_labels = dev.upload_labels(labels, tile.device)
            self.process_tile_device(tilecode, tile, d_labels)
            labels = np.array(labels, copy=True)
            dev.download_labels(d_labels, labels)
            return labels
        points = np.vstack((pointcloud.x, pointcloud.y, pointcloud.z)).T
        return self.process_cloud(tilecode, points, labels, mask)

    def _save_file(self, pointcloud, labels, out_file, start):
        las_utils.label_and_save_las(pointcloud, labels, out_file)
        duration = time.time() - start
        stats = analysis_tools.get_label_stats(labels)
        logger.info('\nSTATISTICS\n' + stats)
        logger.info(f'File processed in {duration:.2f}s, ' + f'output written to {out_file}.\n' + '=' * 20)

    def process_file(self, in_file, out_file=None, mask=None):
        """Process a single LAS / LAZ file and save the result (pipeline.py:99-137); ``.laz`` goes through laspy when it is
        installed, else through the native LASzip decoder / encoder (point formats 0-3)."""
        logger.info(f'Processing file {in_file}.')
        start = time.time()
        if not os.path.isfile(in_file):
            logger.error('The input file specified does not exist')
            return None
        if out_file is None:
            out_file = in_file
        tilecode, pointcloud, labels = self._load_file(in_file)
        labels = self._label_file(tilecode, pointcloud, labels, mask)
        self._save_file(pointcloud, labels, out_file, start)

    def process_folder(self, in_folder, out_folder=None, in_prefix='', out_prefix='', suffix='',
                       hide_progress=False, workers=1):
        """Process a folder of LAS files and save each processed file.

        ``workers`` (not in the reference): with 1, file ``i + 1`` is read and file ``i - 1`` written on
        host threads while the device labels file ``i``; with more, that many files are in flight, each
        read, labelled (own CUDA stream) and written by its own host thread, as in ``process_clouds``."""
        if not os.path.isdir(in_folder):
            logger.error('The input path specified does not exist')
            return None
        in_folder = pathlib.Path(in_folder)
        if out_folder is None:
            out_folder = in_folder
        else:
            pathlib.Path(out_folder).mkdir(parents=True, exist_ok=True)
        if suffix is None:
            suffix = ''
        logger.info('\n===== PIPELINE =====\n' + f'Processing folder {in_folder}, ' +
                    f'writing results in {out_folder}.')
        files = [f for f in in_folder.glob('*')
                 if f.name.endswith(self.FILE_TYPES) and f.name.startswith(in_prefix)]
        logger.debug(f'{len(files)} files found.')
        jobs = []
        for file in files:
            filename, extension = os.path.splitext(file.name)
            if in_prefix and out_prefix:
                filename = filename.replace(in_prefix, out_prefix)
            elif out_prefix:
                filename = out_prefix + filename
            jobs.append((file.as_posix(), os.path.join(out_folder, filename + suffix + extension)))
        workers = max(1, min(int(workers), len(jobs)))
        if workers > 1 and all(hasattr(obj, '_apply_device') for obj in self.processors):
            def body(i):
                in_file, outfile = jobs[i]
                logger.info(f'Processing file {in_file}.')
                start = time.time()
                tilecode, pointcloud, labels = self._load_file(in_file)
                labels = self._label_file(tilecode, pointcloud, labels)
                self._save_file(pointcloud, labels, outfile, start)
            self._run_tile_workers(len(jobs), workers, body)
            logger.info(f'Pipeline finished, {len(files)} processed.\n' + '=' * 20)
            return
        # Files are independent (pipeline.py:186-194 handles them one after the other): while the device
        # labels file i, one host thread already reads and parses file i+1 and another writes file i-1.
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=1) as reader, ThreadPoolExecutor(max_workers=1) as writer:
            loading = reader.submit(self._load_file, jobs[0][0]) if jobs else None
            saving = None
            for i, (in_file, outfile) in enumerate(jobs):
                logger.info(f'Processing file {in_file}.')
                start = time.time()
                tilecode, pointcloud, labels = loading.result()
                loading = reader.submit(self._load_file, jobs[i + 1][0]) if i + 1 < len(jobs) else None
                labels = self._label_file(tilecode, pointcloud, labels)
                if saving is not None:
                    saving.result()                         # one write in flight at a time; surfaces I/O errors
                saving = writer.submit(self._save_file, pointcloud, labels, outfile, start)
            if saving is not None:
                saving.result()
        logger.info(f'Pipeline finished, {len(files)} processed.\n' + '=' * 20)

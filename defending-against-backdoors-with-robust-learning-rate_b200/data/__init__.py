"""Data layer: device-resident datasets, the reference's partitioner and the backdoor poisoner.

Reference counterparts: ``H5Dataset`` / ``DatasetSplit`` (src/utils.py:11-54), ``distribute_data``
(src/utils.py:58-92), ``get_datasets`` (src/utils.py:95-124), ``poison_dataset`` / ``add_pattern_bd``
(src/utils.py:160-284).
"""
from .datasets import DeviceDataset, DatasetSplit, H5Dataset, get_datasets, make_synthetic, DATASET_META
from .partition import distribute_data
from .poison import poison_dataset, add_pattern_bd, pattern_pixels, make_poisoned_val

__all__ = ["DeviceDataset", "DatasetSplit", "H5Dataset", "get_datasets", "make_synthetic", "DATASET_META",
           "distribute_data", "poison_dataset", "add_pattern_bd", "pattern_pixels", "make_poisoned_val"]

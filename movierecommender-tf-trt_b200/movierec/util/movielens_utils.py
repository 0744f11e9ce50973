""" Movielens constants and ratings loader (reference movierec/util/movielens_utils.py).

Only what the train/eval hot path needs is here: dataset constants (reference :16-55) and
`load_ratings_data` (:104-125).  Downloading (:128-170) and the movies table (:85-101) are outside
the hot path (no network in the build environment) -- a missing file raises FileNotFoundError.
"""

import os

import numpy as np
import pandas as pd

ML_100K = 'ml-100k'
ML_1M = 'ml-1m'
ML_20M = 'ml-20m'
MOVIELENS_DATASET_NAMES = [ML_100K, ML_1M, ML_20M]

RATINGS_FILE_NAME = {ML_100K: 'u.data', ML_1M: 'ratings.dat', ML_20M: 'ratings.csv'}
SEPARATOR = {ML_100K: '\t|\\|', ML_1M: '::', ML_20M: ','}
HAS_HEADER = {ML_100K: False, ML_1M: False, ML_20M: True}

# Table sizes the reference hard-codes (max-id based, :45-55).
NUM_USERS = {ML_100K: 943, ML_1M: 6040, ML_20M: 138493}
NUM_ITEMS = {ML_100K: 1682, ML_1M: 3952, ML_20M: 27278}


def get_path(data_dir, dataset_name, file_name):
    return os.path.join(data_dir, dataset_name, file_name)


def get_ratings_path(data_dir, dataset_name):
    return get_path(data_dir, dataset_name, RATINGS_FILE_NAME[dataset_name])


def download_movielens(dataset_name, output_dir):
    raise FileNotFoundError('Downloading {} is not supported in this build (no network); place the ratings '
                            'file under {}'.format(dataset_name, os.path.join(output_dir, dataset_name)))


def load_ratings_data(data_dir, dataset_name, col_user_id='userId', col_item_id='itemId', col_rating='rating',
                      download=True):
    """Ratings as a DataFrame with 0-based int32 ids and float32 ratings (reference :104-125)."""
    path = get_ratings_path(data_dir, dataset_name)
    if not os.path.exists(path):
        if download:
            download_movielens(dataset_name, data_dir)
        raise FileNotFoundError('{} not found. Download the dataset first or set param download=True.'.format(path))
    df = pd.read_csv(path, sep=SEPARATOR[dataset_name], header=0 if HAS_HEADER[dataset_name] else None,
                     encoding='utf-8', engine='python', usecols=(0, 1, 2), names=(col_user_id, col_item_id, col_rating),
                     dtype={col_user_id: np.int32, col_item_id: np.int32, col_rating: np.float32})
    df[col_user_id] = df[col_user_id] - 1  # ids are 1-based in the files
    df[col_item_id] = df[col_item_id] - 1
    return df

"""Training entry point with the reference's signature and CLI (movierec/trainer.py:30, :86-96).

    train(model_name, dataset_name, data_dir, output_dir, params=DEFAULT_PARAMS, verbose=1)
    python -m movierec.trainer -m NAME -n ml-100k [-d data/] [-o models] [-l INFO]

Wiring (reference :52-80): split the ratings leave-last-two-out, a shuffled train generator that
ignores validation/test when sampling negatives, an unshuffled validation generator that excludes
the train positives, table sizes taken from the generator, fit with the model's callbacks, save.
"""

import argparse
import logging

try:
    from . import data_pipeline
    from .model import MovierecModel
except ImportError:  # pragma: no cover -- `python trainer.py` from inside the package directory
    import data_pipeline
    from model import MovierecModel

# Hyper-parameters of the reference run (trainer.py:8-27); num_users / num_items are filled in by
# train() from the dataset, which mutates the dict it is given exactly as the reference does.
DEFAULT_PARAMS = dict(
    layers_sizes=[64, 32, 16, 8], layers_l2reg=[0, 0, 0, 0],
    optimizer="adam", lr=0.001, beta_1=0.9, beta_2=0.999,
    batch_size=100, batch_size_eval=200, num_negs_per_pos=9, num_negs_per_pos_eval=99, k=5, epochs=20,
)


def build_generators(dataset_name, train_df, validation_df, params):
    """(train generator, validation generator) as reference trainer.py:55-69."""
    make = data_pipeline.MovieLensDataGenerator
    fit_gen = make(dataset_name, train_df, params["batch_size"], params["num_negs_per_pos"],
                   extra_data_df=None, shuffle=True)
    val_gen = make(dataset_name, validation_df, params["batch_size_eval"], params["num_negs_per_pos_eval"],
                   extra_data_df=train_df, shuffle=False)
    return fit_gen, val_gen


def train(model_name, dataset_name, data_dir, output_dir, params=DEFAULT_PARAMS, verbose=1):
    """Train a model on a MovieLens dataset and save it under `output_dir`; returns the model
    (the reference returns nothing)."""
    splits = data_pipeline.load_ratings_train_test_sets(dataset_name, data_dir)
    fit_gen, val_gen = build_generators(dataset_name, splits[0], splits[1], params)
    params["num_users"], params["num_items"] = fit_gen.num_users, fit_gen.num_items

    recommender = MovierecModel(params, model_name, output_dir, verbose)
    recommender.log_summary()
    recommender.fit_generator(fit_gen, val_gen, params["epochs"])
    recommender.save()
    return recommender


def main(argv=None):
    cli = argparse.ArgumentParser(description="Train a NeuMF recommender on a B200.")
    cli.add_argument("-m", "--model-name", required=True, help="name used for the saved files")
    cli.add_argument("-n", "--dataset-name", required=True, help="ml-100k | ml-1m | ml-20m")
    cli.add_argument("-d", "--data-dir", default="data/", help="directory holding <dataset>/<ratings file>")
    cli.add_argument("-o", "--output-dir", default="models", help="where weights and params are written")
    cli.add_argument("-l", "--log-level", default="INFO")
    opts = cli.parse_args(argv)
    level = logging.getLevelName(opts.log_level)
    logging.getLogger().setLevel(level)
    logging.info("training with %s", DEFAULT_PARAMS)
    return train(opts.model_name, opts.dataset_name, opts.data_dir, opts.output_dir, DEFAULT_PARAMS, level)


if __name__ == "__main__":
    main()

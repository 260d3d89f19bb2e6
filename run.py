#!/usr/bin/env python3
"""Run shoeprint image retrieval: same entry point, same ``run.toml`` as the reference ``run.py``."""

from src.shoeprint_image_retrieval.config import load_config
from src.shoeprint_image_retrieval.dataloader import Dataloader
from src.shoeprint_image_retrieval.network import Model
from src.shoeprint_image_retrieval.parse_results import cmp_all
from src.shoeprint_image_retrieval.similarity import compare_maps


def main(config_path: str = "run.toml") -> None:
    config = load_config(config_path)
    dataloader = Dataloader(config)
    print(f"{dataloader.num_clusters} clusters of image sizes found.")

    for shoemark_images, shoeprint_images, matching_shoeprint_ids, block in dataloader:
        print(f"Cluster has {len(shoemark_images)} items.")
        model = Model(config, block)
        shoemark_features = model.get_multiple_feature_maps(shoemark_images)
        shoeprint_features = model.get_multiple_feature_maps(shoeprint_images)
        print("Calculating ranks:")
        ranks = compare_maps(shoemark_features, shoeprint_features, matching_shoeprint_ids, config)
        cmp_all(
            list(ranks),
            total_shoeprints=len(dataloader.shoeprint_files),
            total_shoemarks=len(dataloader.shoemark_files),
        )


if __name__ == "__main__":
    main()

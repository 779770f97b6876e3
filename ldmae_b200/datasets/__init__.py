from .img_latent_dataset import ImgLatentDataset, write_latent_shard  # noqa: F401

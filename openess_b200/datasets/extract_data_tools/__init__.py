"""Mirror of datasets/extract_data_tools/ (DDD17 recording access)."""

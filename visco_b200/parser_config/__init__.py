"""Command line: ``visco compressms ...`` / ``visco decompressms ...`` — same group, subcommand names, option names,
single-dash abbreviations and defaults as the reference (visco/parser_config/__init__.py:5-14, compressms.yaml,
decompressms.yaml). The reference generates its click options from the YAML files with stimela/scabha; neither is
installed here, so the same table is expressed with plain click."""
import click

import visco_b200

# (name, abbreviation, type, default, required, help)   -- reference visco/parser_config/compressms.yaml:2-168
COMPRESS_OPTIONS = [
    ("ms", "ms", str, None, True, "The Measurement Set file path."),
    ("zarrstore", "zs", str, None, True, "The path to the output Zarr store."),
    ("consolidated", "consol", bool, True, False, "Consolidate metadata."),
    ("chunk_size_row", "csr", int, 10000, False, "Row chunk size."),
    ("overwrite", "ow", bool, True, False, "Overwrite an existing store."),
    ("compressor", None, str, "zstd", False, "Codec of the factor arrays: zstd, gzip, blosc."),
    ("level", "l", int, 4, False, "Codec level."),
    ("nworkers", "nw", int, 4, False, "Dask workers (accepted, ignored: batches run on the GPU)."),
    ("nthreads", "nt", int, 2, False, "Threads per worker (accepted, ignored)."),
    ("memory_limit", "ml", str, "4GB", False, "Worker memory limit (accepted, ignored)."),
    ("direct_to_workers", "dtw", bool, True, False, "(accepted, ignored)"),
    ("dashboard_address", "da", str, None, False, "(accepted, ignored)"),
    ("host_address", "ha", str, None, False, "(accepted, ignored)"),
    ("correlation", "corr", str, "XX,YY", False, "Correlations to compress."),
    ("correlation_optimized", "co", bool, False, False, "Stack XX+YY and XY+YX before the SVD."),
    ("fieldid", "fid", int, 0, False, "FIELD_ID."),
    ("ddid", None, int, 0, False, "DATA_DESC_ID."),
    ("scan", None, int, 1, False, "SCAN_NUMBER."),
    ("column", "col", str, "DATA", False, "Column to compress."),
    ("outcolumn", "oc", str, "COMPRESSED_DATA", False, "Name of the compressed column in the store."),
    ("batch_size", "bs", int, 20, False, "Baselines per batch."),
    ("use_model_data", "umd", bool, None, False, "Replace flagged data with model data."),
    ("model_data", "md", str, None, False, "Model data column."),
    ("flagestimate", "fest", bool, None, False, "Estimate flagged values with scipy griddata (out of scope of this build: raises)."),
    ("flagvalue", "fv", str, None, False, "Constant that replaces flagged values, e.g. 0 or 1+1j."),
    ("decorrelation", "dec", float, None, False, "Keep singular values up to this decorrelation (energy = dec^2)."),
    ("compressionrank", "cr", int, None, False, "Keep this many singular values (wins over --decorrelation)."),
    ("antennas", None, str, None, False, "List of antenna indices, e.g. '[0,1,2]'."),
    # additive (not in the reference): GPUs of this box to spread the baselines over
    ("ngpus", "ng", int, 1, False, "GPUs to shard the baselines over (one host thread and handle per GPU)."),
]
# reference visco/parser_config/decompressms.yaml:2-26
DECOMPRESS_OPTIONS = [
    ("zarrstore", "zs", str, None, True, "Path to the zarr store with the compressed data components."),
    ("ms", "ms", str, "decompressed.ms", False, "The output Measurement Set."),
    ("column", "col", str, "COMPRESSED_DATA", False, "Compressed column to decompress."),
    ("batch_size", "bs", int, 50, False, "Reconstruction tasks per batch."),
    ("ngpus", "ng", int, 1, False, "GPUs to shard the reconstruction tasks over (additive option)."),
]


def _clickify(options):
    def deco(f):
        for name, abbr, typ, default, required, helptext in reversed(options):
            long = "--" + name.replace("_", "-")
            decls = [long] + ([f"-{abbr}"] if abbr else []) + [name]
            if typ is bool:
                f = click.option(*[f"{long}/--no-{name.replace('_', '-')}"] + ([f"-{abbr}"] if abbr else []) + [name],
                                 default=default, help=helptext)(f)
            else:
                f = click.option(*decls, type=typ, default=default, required=required, help=helptext, show_default=True)(f)
        return f
    return deco


@click.group(help="A tool for compressing radio interferometric data using lossy Singular Value Decomposition (SVD) "
                  "techniques.\n\nAlso includes utilities for decompressing the data back to a Measurement Set (MS) "
                  "format.\n\nB200 build: the SVD / truncation / reconstruction run on the GPU (libvisco_b200).")
@click.version_option(str(visco_b200.__version__))
def cli():
    pass


@cli.command("compressms")
@click.version_option(str(visco_b200.__version__))
@_clickify(COMPRESS_OPTIONS)
def compressrunit(**kw):
    """Compress a Measurement Set (reference parser_config/compressms.py:23-86)."""
    import ast

    from visco_b200 import compress_ms
    antennas = kw["antennas"]
    if isinstance(antennas, str):
        try:
            antennas = ast.literal_eval(antennas)
        except (ValueError, SyntaxError):
            raise click.BadParameter(f"Invalid format for antennas: {antennas}")
    compress_ms.compress_full_ms(
        ms_path=kw["ms"], zarr_path=kw["zarrstore"], consolidated=kw["consolidated"], chunk_size_row=kw["chunk_size_row"],
        overwrite=kw["overwrite"], compressor=kw["compressor"], nworkers=kw["nworkers"], nthreads=kw["nthreads"],
        memory_limit=kw["memory_limit"], direct_to_workers=kw["direct_to_workers"], level=kw["level"],
        correlation=kw["correlation"], correlation_optimized=kw["correlation_optimized"], fieldid=kw["fieldid"],
        ddid=kw["ddid"], scan=kw["scan"], column=kw["column"], outcolumn=kw["outcolumn"], batch_size=kw["batch_size"],
        dashboard_addr=kw["dashboard_address"], host_addr=kw["host_address"], use_model_data=bool(kw["use_model_data"]),
        model_data=kw["model_data"], flag_estimate=bool(kw["flagestimate"]), decorrelation=kw["decorrelation"],
        compressionrank=kw["compressionrank"], flagvalue=kw["flagvalue"], antennas=antennas, ngpus=kw["ngpus"])


@cli.command("decompressms")
@click.version_option(str(visco_b200.__version__))
@_clickify(DECOMPRESS_OPTIONS)
def decompressrunit(**kw):
    """Decompress a store back to a Measurement Set (reference parser_config/decompressms.py:23-33)."""
    from visco_b200 import decompress_ms
    decompress_ms.write_datasets_to_ms(kw["zarrstore"], kw["ms"], kw["column"], kw["batch_size"], ngpus=kw["ngpus"])


def main():
    cli()

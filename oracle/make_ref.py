"""Recipe for oracle/_ref: a verbatim, git-ignored copy of the reference's Python import closure for the render_rays path.

TEST / BASELINE INFRASTRUCTURE ONLY.  Run in the build container (the only place /root/reference exists):

    python oracle/make_ref.py          # also run by __graft_entry__.build() when /root/reference is present

It imports the UNMODIFIED reference renderer (NeRFs/HeadNeRF/train/audio_exp_nerf.py) and the torso helpers through
oracle/ref_import.py, looks at which files under /root/reference the interpreter actually loaded, and copies exactly those files
into oracle/_ref/ with their relative paths.  oracle/_ref/ is listed in .gitignore (reference sources never enter the history) but
not in .gpurunignore, so it travels to the GPU box, where `bench.py --impl reference` and the `cpu_baseline` leg time the
reference's own Network.render_rays instead of the oracle port.  Nothing in ideal-nerf_b200/ reads it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")


def main():
    if not os.path.isdir(SRC):
        print(f"{SRC} not present: keeping whatever oracle/_ref already holds")
        return 0
    sys.path.insert(0, ROOT)
    os.environ["IDEAL_NERF_REFERENCE"] = SRC
    from oracle import ref_import
    ref_import.import_head(force_cpu=True)
    ref_import.import_torso_helpers()
    files = sorted({os.path.realpath(m.__file__) for m in list(sys.modules.values())
                    if getattr(m, "__file__", None) and os.path.realpath(m.__file__).startswith(SRC + os.sep)})
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for f in files:
        rel = os.path.relpath(f, SRC)
        os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
        shutil.copyfile(f, os.path.join(DST, rel))
    with open(os.path.join(DST, "MANIFEST.txt"), "w") as fh:
        fh.write("verbatim copies of /root/reference files (import closure of the render_rays path); made by oracle/make_ref.py\n")
        fh.writelines(os.path.relpath(f, SRC) + "\n" for f in files)
    print(f"oracle/_ref: {len(files)} files")
    return 0


if __name__ == "__main__":
    sys.exit(main())

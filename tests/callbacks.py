"""Callback sources of the reference's test/callback.cpp (:55-63 load, :151-158 store), kept in
the reference's OpenCL-C form (the library translates the subset) plus CUDA C++ equivalents."""


def load_zero_pad_opencl(real, M, N_ext_spec, N_spec):
    """c2r load callback: the stored spectrum has N_spec rows, the transform expects N_ext_spec;
    rows beyond N_spec read as zero."""
    return """
%s2 load(global %s2* in, size_t offset) {
    uint n = offset / %d %% %d;
    if (n < %d) {
        uint k = offset / %d;
        return in[offset - k * %d];
    }
    return 0;
}""" % (real, real, M, N_ext_spec, N_spec, M * N_ext_spec, M * (N_ext_spec - N_spec))


def load_zero_pad_cuda(real, M, N_ext_spec, N_spec):
    return """
__device__ %s2 load(%s2 const* in, size_t offset) {
    unsigned n = offset / %d %% %d;
    if (n < %d) {
        unsigned k = offset / %d;
        return in[offset - k * %d];
    }
    %s2 z; z.x = 0; z.y = 0;
    return z;
}""" % (real, real, M, N_ext_spec, N_spec, M * N_ext_spec, M * (N_ext_spec - N_spec), real)


def store_truncate_scale_opencl(real, M, N_spec, N_cut, scale):
    """r2c store callback: keep the first N_cut rows of the spectrum, scaled by `scale`."""
    return """
void store(global %s2* out, size_t offset, %s2 value) {
    uint n = offset / %d %% %d;
    if (n < %d) {
        uint k = offset / %d;
        out[offset - k * %d] = value * ((%s) %s);
    }
}""" % (real, real, M, N_spec, N_cut, M * N_spec, M * (N_spec - N_cut), real, float(scale).hex())

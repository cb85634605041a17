// TEST INFRASTRUCTURE ONLY: the reference's PathAligner.cpp includes common/Alignment.hh without using anything from
// it; the real header pulls in htslib (variant/RefVar.hh -> common/BCFHelpers.hh), which is not in this image.
#pragma once

// TEST INFRASTRUCTURE ONLY: the reference's KmerAligner.cpp includes common/Klib.hh without using anything from it;
// the real header derives from common/Alignment.hh, which pulls in htslib (not in this image).
#pragma once

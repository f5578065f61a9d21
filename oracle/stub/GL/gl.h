/* Stub for the unused <cuda_gl_interop.h> include at ref simulator.cu:1 --
 * this image has no OpenGL headers.  Test infrastructure only. */
#pragma once
typedef unsigned int GLuint;
typedef unsigned int GLenum;

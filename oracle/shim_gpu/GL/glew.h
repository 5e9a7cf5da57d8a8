// Headless stand-in for <GL/glew.h>: the handful of buffer-object calls the reference's
// gpu/src/particlesystem.cpp makes (lines 95,129,316-341,708-722) are forwarded to a plain
// cudaMalloc'd buffer implemented in oracle/ref_gpu_glue.cu.  Oracle build only.
#pragma once
#include "gl.h"
#define GL_ARRAY_BUFFER 0x8892
#define GL_DYNAMIC_DRAW 0x88E8
extern "C" {
void glGenBuffers(GLsizei n, GLuint *ids);
void glDeleteBuffers(GLsizei n, const GLuint *ids);
void glBindBuffer(GLenum target, GLuint id);
void glBufferData(GLenum target, GLsizeiptr size, const void *data, GLenum usage);
void glBufferSubData(GLenum target, GLintptr off, GLsizeiptr size, const void *data);
}

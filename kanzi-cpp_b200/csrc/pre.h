// pre.h -- host stages of the block pipeline (pre.cu, pretext.cu): interfaces used by api.cu.
#pragma once
#include <math.h>

#include "common.cuh"

// wire-visible transform ids of the host stages (transform/TransformFactory.hpp:49-73)
#define KNZ_T_TEXT 10
#define KNZ_T_MM 15
#define KNZ_T_UTF 17
#define KNZ_T_PACK 18
#define KNZ_T_DNA 19

// Global::DataType (Global.hpp:29): what a stage learnt about the block, handed to the stages behind it
enum { KDT_UNDEFINED = 0, KDT_TEXT, KDT_MULTIMEDIA, KDT_EXE, KDT_NUMERIC, KDT_BASE64, KDT_DNA, KDT_BIN, KDT_UTF8, KDT_SMALL_ALPHABET };

// The entries of the reference's Context the host stages read or write
struct KnzPreCtx {
    int dataType;  // "dataType"
    int blockSize; // "blockSize" (TextCodec sizes its hash table by it)
    int eType;     // entropy id of the stream (selects the text codec variant, TransformFactory.hpp:227-242)
};

bool knz_is_host_stage(int type);
int knz_pre_max_len(int type, int n); // Transform::getMaxEncodedLength
// Transform<byte>::forward / inverse of one stage: true = applied (*outLen bytes in dst), false = refused
bool knz_pre_forward(int type, const u8* src, int n, u8* dst, int cap, int* outLen, KnzPreCtx* pc);
bool knz_pre_inverse(int type, const u8* src, int n, u8* dst, int cap, int* outLen, const KnzPreCtx* pc);
int knz_detect_simple_type(int n, const u32 f[256]);
bool knz_utf8_plausible(const u32 f0[256], const u32* f1, int n);
void knz_log2_table(int tab[257]);
int knz_magic_data_type(const u8* block, int n);
bool knz_magic_known(const u8* block); // Magic::getType(block) != NO_MAGIC (4 readable bytes)
// pretext.cu
bool knz_text_forward(const u8* src, int n, u8* dst, int cap, int* outLen, KnzPreCtx* pc);
bool knz_text_inverse(const u8* src, int n, u8* dst, int cap, int* outLen, int blockSize, int eType);
// entropy ids as the stages need them
#ifndef E_RAW
#define E_RAW 0
#define E_HUF 1
#define E_ANS0 5
#endif

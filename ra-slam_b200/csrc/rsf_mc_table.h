// rsf_mc_table.h -- the classic 256-case marching-cubes triangle table (Lorensen & Cline 1987; the widely
// reproduced public-domain table of P. Bourke, "Polygonising a scalar field"), in the corner / edge numbering
// the reference uses (utils/tsdf/mcube_table.cuh:14-31, :75).  Packed: nibble k of word c is the k-th edge
// index (0..11) of case c, 0xf terminates the list; three consecutive nibbles form one triangle.
// Sanity properties checked when the packing was generated: 256 rows, triangle-count histogram
// {0:2, 1:16, 2:50, 3:80, 4:76, 5:32}, case c and 255-c use the same edge set.
#pragma once
#define RSF_MC_TRI_TABLE \
  0xfffffffffffffffull, 0xffffffffffff380ull, 0xffffffffffff910ull, 0xfffffffff189381ull, \
  0xffffffffffffa21ull, 0xfffffffffa21380ull, 0xfffffffff920a29ull, 0xffffff89a8a2382ull, \
  0xffffffffffff2b3ull, 0xfffffffff0b82b0ull, 0xfffffffffb32091ull, 0xffffffb89b912b1ull, \
  0xfffffffff3ab1a3ull, 0xffffffab8a801a0ull, 0xffffff9ab9b3093ull, 0xfffffffffb8aa89ull, \
  0xffffffffffff874ull, 0xfffffffff437034ull, 0xfffffffff748910ull, 0xffffff137174914ull, \
  0xfffffffff748a21ull, 0xffffffa21403743ull, 0xffffff748209a29ull, 0xfff4973727929a2ull, \
  0xfffffffff2b3748ull, 0xffffff40242b74bull, 0xffffffb32748109ull, 0xfff1292b9b49b74ull, \
  0xffffff487ab31a3ull, 0xfff4b7401b41ab1ull, 0xfff30bab9b09874ull, 0xffffffab99b4b74ull, \
  0xffffffffffff459ull, 0xfffffffff380459ull, 0xfffffffff051450ull, 0xffffff513538458ull, \
  0xfffffffff459a21ull, 0xffffff594a21803ull, 0xffffff204245a25ull, 0xfff8434535235a2ull, \
  0xfffffffffb32459ull, 0xffffff594b802b0ull, 0xffffffb32510450ull, 0xfff584b82852512ull, \
  0xffffff45931ab3aull, 0xfffab81a8180594ull, 0xfff30bab5b05045ull, 0xffffffb8aa85845ull, \
  0xfffffffff975879ull, 0xffffff375359039ull, 0xffffff751710870ull, 0xfffffffff753351ull, \
  0xffffff21a759879ull, 0xfff37503505921aull, 0xfff25a758528208ull, 0xffffff7533525a2ull, \
  0xffffff2b3987597ull, 0xfffb72029279759ull, 0xfff751871810b32ull, 0xffffff51771b12bull, \
  0xfffb3a31a758859ull, 0x0aba010b7905075ull, 0x07570805a30b0abull, 0xfffffffff5b75abull, \
  0xffffffffffff56aull, 0xfffffffff6a5380ull, 0xfffffffff6a5109ull, 0xffffff6a5891381ull, \
  0xfffffffff162561ull, 0xffffff803621561ull, 0xffffff620609569ull, 0xfff823625285895ull, \
  0xfffffffff56ab32ull, 0xffffff56a02b80bull, 0xffffff6a5b32910ull, 0xfffb892b92916a5ull, \
  0xffffff315356b36ull, 0xfff6b51505b0b80ull, 0xfff9505606306b3ull, 0xffffff89bb96956ull, \
  0xfffffffff8746a5ull, 0xffffffa56374034ull, 0xffffff7486a5091ull, 0xfff49737179156aull, \
  0xffffff874156216ull, 0xfff743403625521ull, 0xfff620560509748ull, 0x962695923497937ull, \
  0xffffff56a4872b3ull, 0xfffb720242746a5ull, 0xfff6a5b32874910ull, 0x6a54b7b492b9129ull, \
  0xfff6b51535b3748ull, 0xb404b7b016b5b15ull, 0x74836b630560950ull, 0xfff9b7974b96956ull, \
  0xfffffffffa4694aull, 0xffffff380a946a4ull, 0xffffff04606a10aull, 0xfffa16468618138ull, \
  0xffffff462421941ull, 0xfff462942921803ull, 0xfffffffff624420ull, 0xffffff624428238ull, \
  0xffffff32b46a94aull, 0xfff6a4a94b82280ull, 0xfffa164606102b3ull, 0x1b8b12184a16146ull, \
  0xfff36b319639469ull, 0x14641916b0181b8ull, 0xffffff4600636b3ull, 0xfffffffff86b846ull, \
  0xffffffa98a876a7ull, 0xfffa76a907a0370ull, 0xfff0818717a176aull, 0xffffff37117a76aull, \
  0xfff768981861621ull, 0x937390976192962ull, 0xffffff206607087ull, 0xfffffffff276237ull, \
  0xfff76898a86ab32ull, 0x7a9a76790b72702ull, 0xb32a767a1871081ull, 0xfff17616a71b12bull, \
  0x63136b619768698ull, 0xfffffffff76b190ull, 0xfff06b0b3607087ull, 0xffffffffffff6b7ull, \
  0xffffffffffffb67ull, 0xfffffffff67b803ull, 0xfffffffff67b910ull, 0xffffff67b138918ull, \
  0xfffffffff7b621aull, 0xffffff7b6803a21ull, 0xffffff7b69a2092ull, 0xfff89a38a3a27b6ull, \
  0xfffffffff726327ull, 0xffffff026067807ull, 0xffffff910732672ull, 0xfff678891681261ull, \
  0xffffff73171a67aull, 0xfff801781a7167aull, 0xfff7a69a0a70730ull, 0xffffff9a88a7a67ull, \
  0xfffffffff68b486ull, 0xffffff640603b63ull, 0xffffff109648b68ull, 0xfff63b139369649ull, \
  0xffffff1a28b6486ull, 0xfff640b60b03a21ull, 0xfff9a2920b648b4ull, 0x36463b34923a39aull, \
  0xffffff264248328ull, 0xfffffffff264240ull, 0xfff834642432091ull, 0xffffff642241491ull, \
  0xfff1a6648168318ull, 0xffffff40660a01aull, 0x39a9303a6834364ull, 0xfffffffff4a649aull, \
  0xfffffffffb67594ull, 0xffffff67b594380ull, 0xffffffb67045105ull, 0xfff51345343867bull, \
  0xffffffb6721a459ull, 0xfff594380a217b6ull, 0xfff204a24a45b67ull, 0x67b25a523453843ull, \
  0xffffff945267327ull, 0xfff786260680459ull, 0xfff045051673263ull, 0x851584812786826ull, \
  0xfff73167161a459ull, 0x459078701671a61ull, 0xa737a6a305a4a04ull, 0xfffa84a458a7a67ull, \
  0xffffff98b9b6596ull, 0xfff590650360b63ull, 0xfffb65510b508b0ull, 0xffffff1355363b6ull, \
  0xfff65b8b9b59a21ull, 0xa21965690b603b0ull, 0x52025a50865b58bull, 0xfff35a3a25363b6ull, \
  0xfff283265825985ull, 0xffffff260069659ull, 0x826283865081851ull, 0xfffffffff612651ull, \
  0x698965683a61631ull, 0xfff06505960a01aull, 0xfffffffffa65830ull, 0xffffffffffff65aull, \
  0xfffffffffb57a5bull, 0xffffff03857ba5bull, 0xffffff091ba57b5ull, 0xfff1381897ba57aull, \
  0xffffff15717b21bull, 0xfffb27571721380ull, 0xfff7b2209729579ull, 0x289823295b27257ull, \
  0xffffff573532a52ull, 0xfff52a578258028ull, 0xfff2a37353a5109ull, 0x25752a278129289ull, \
  0xfffffffff573531ull, 0xffffff571170780ull, 0xffffff735539309ull, 0xfffffffff795789ull, \
  0xffffff8ba8a5485ull, 0xfff03bba50b5405ull, 0xfff54aba8a48910ull, 0x41314943b54a4baull, \
  0xfff8548b2582152ull, 0xb151b2b543b0b40ull, 0x58b8545b2950520ull, 0xfffffffff3b2549ull, \
  0xfff483543253a52ull, 0xffffff0244252a5ull, 0x910854583a532a3ull, 0xfff2492914252a5ull, \
  0xffffff153358548ull, 0xfffffffff501540ull, 0xfff530509358548ull, 0xffffffffffff549ull, \
  0xffffffba9b947b4ull, 0xfffba97b9794380ull, 0xfffb470414b1ba1ull, 0x4bab474a1843413ull, \
  0xfff219b294b97b4ull, 0x3801b2b197b9479ull, 0xffffff04224b47bull, 0xfff42343824b47bull, \
  0xfff947732972a92ull, 0x70207872a4797a9ull, 0xa040a1a472a3a73ull, 0xfffffffff4782a1ull, \
  0xffffff317714194ull, 0xfff178180714194ull, 0xfffffffff347304ull, 0xffffffffffff784ull, \
  0xfffffffff8ba8a9ull, 0xffffffa9bb93903ull, 0xffffffba88a0a10ull, 0xfffffffffa3ba13ull, \
  0xffffff8b99b1b21ull, 0xfff9b2921b93903ull, 0xfffffffffb08b20ull, 0xffffffffffffb23ull, \
  0xffffff98aa82832ull, 0xfffffffff2902a9ull, 0xfff8a1810a82832ull, 0xffffffffffff2a1ull, \
  0xfffffffff819831ull, 0xffffffffffff190ull, 0xffffffffffff830ull, 0xfffffffffffffffull,
